// A smooth non-convex test objective with a failure region, shared by the driver that runs the
// REFERENCE optimisers (oracle/ref/ref_scg.cpp) and the one that runs ours
// (tests/cpp/host_check.cpp), so both see bit-identical function values.
#ifndef ANALYTIC_OBJECTIVE_H
#define ANALYTIC_OBJECTIVE_H
#include <cmath>
#include <vector>

// returns false ("evaluation failed") when any |x_i| > 6: exercises the step-halving branch
inline bool analytic_objective(const std::vector<double> &x, double &f, std::vector<double> &g)
{
    const size_t n = x.size();
    for (size_t i = 0; i < n; i++)
        if (std::fabs(x[i]) > 6.0) return false;
    f = 0.0;
    g.assign(n, 0.0);
    for (size_t i = 0; i < n; i++) {
        const double c = 0.3 * std::sin(1.7 * (double)i) + 0.5, w = 1.0 + 0.15 * (double)i;
        const double d = x[i] - c;
        f += w * d * d + 0.05 * d * d * d * d;
        g[i] += 2.0 * w * d + 0.2 * d * d * d;
        if (i + 1 < n) {
            const double p = x[i] * x[i + 1];
            f += 0.5 * std::sin(p);
            g[i] += 0.5 * std::cos(p) * x[i + 1];
            g[i + 1] += 0.5 * std::cos(p) * x[i];
        }
    }
    return true;
}
#endif
