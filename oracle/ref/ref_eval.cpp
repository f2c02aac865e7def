// ref_eval -- TEST/BENCH INFRASTRUCTURE: drives the UNMODIFIED reference classes
// (compiled from /root/reference by oracle/ref/build_ref.sh) at the drop-in boundary
// c_objective_one::compute_objective / GP_Regression::predict
// (medgpc/src/util/c_objective_one.cpp:40-81, medgpc/src/core/gp_regression.cpp:128-214).
//
// usage: ref_eval <case-file> <mode> <threads> [repeat]
//   mode 0 = NLML only, 1 = NLML + gradient, 2 = predict the listed test points
// case file (text): "Q D R n P" ; n lines "meta x y" ; P lines theta ;
//                   [mode 2: "m" ; m lines "meta x"]
// output: "%.17g" numbers, one per line: ok-flag, nlml, [P gradient entries] or
//         [m means, m variances]; last line "seconds_per_eval <t>".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <chrono>
#include <omp.h>
#include "core/gp_model_include.h"
#include "core/gp_regression.h"
#include "core/c_hyperparam.h"
#include "util/c_objective_one.h"

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: ref_eval <case> <mode> <threads> [repeat]\n"); return 2; }
    int mode = atoi(argv[2]), threads = atoi(argv[3]);
    int repeat = argc > 4 ? atoi(argv[4]) : 1;
    FILE *fp = fopen(argv[1], "r");
    if (!fp) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    int Q, D, R, n, P;
    if (fscanf(fp, "%d %d %d %d %d", &Q, &D, &R, &n, &P) != 5) return 2;
    std::vector<int> meta(n); std::vector<float> x(n), y(n);
    for (int i = 0; i < n; i++) if (fscanf(fp, "%d %f %f", &meta[i], &x[i], &y[i]) != 3) return 2;
    std::vector<double> theta(P);
    for (int i = 0; i < P; i++) if (fscanf(fp, "%lf", &theta[i]) != 1) return 2;
    std::vector<int> meta2; std::vector<float> x2;
    if (mode == 2) {
        int m; if (fscanf(fp, "%d", &m) != 1) return 2;
        meta2.resize(m); x2.resize(m);
        for (int i = 0; i < m; i++) if (fscanf(fp, "%d %f", &meta2[i], &x2[i]) != 2) return 2;
    }
    fclose(fp);

    omp_set_nested(1);
    std::vector<int> kp = {Q, D, R};
    std::vector<int> lp = {D};
    c_kernel_LMC_SM kernel(kp);
    c_inference_prior inffunc(threads);
    c_meanfunc_zero meanfunc;
    c_likelihood_gaussianMO likfunc(lp);
    c_prior prior(Q * (D * R + 2 + D), 0, D);
    c_kernel *kptr = &kernel; c_meanfunc *mptr = &meanfunc; c_likelihood *lptr = &likfunc;
    c_inference *iptr = &inffunc; c_prior *pptr = &prior;

    auto t0 = std::chrono::steady_clock::now();
    if (mode == 0 || mode == 1) {
        c_objective_one obj(7, kp, meta, x, y);
        double f = 0.0; std::vector<double> g; bool ok = false;
        for (int r = 0; r < repeat; r++)
            ok = obj.compute_objective(mode == 1, theta, f, g, kptr, mptr, lptr, iptr, pptr);
        auto t1 = std::chrono::steady_clock::now();
        printf("%d\n%.17g\n", (int)ok, f);
        if (mode == 1) for (size_t i = 0; i < g.size(); i++) printf("%.17g\n", g[i]);
        printf("seconds_per_eval %.9g\n", std::chrono::duration<double>(t1 - t0).count() / repeat);
    } else {
        c_hyperparam hyp(theta, kernel.get_kernel_hyp_num(), 0, D);
        kernel.set_kernel_hyp(hyp.get_hyp_cov());
        meanfunc.set_meanfunc_hyp(hyp.get_hyp_mean());
        likfunc.set_likfunc_hyp(hyp.get_hyp_lik());
        std::vector<std::vector<float> > post;
        bool ok = false;
        for (int r = 0; r < repeat; r++) {
            GP_Regression gpr(1, kptr, mptr, lptr, iptr, pptr);
            gpr.train(false, meta, x, y);
            ok = gpr.get_flag_trained();
            if (ok) post = gpr.predict(meta, meta2, x, y, x2);
        }
        auto t1 = std::chrono::steady_clock::now();
        printf("%d\n%.17g\n", (int)ok, 0.0);
        if (ok) {
            for (size_t i = 0; i < post[0].size(); i++) printf("%.17g\n", (double)post[0][i]);
            for (size_t i = 0; i < post[1].size(); i++) printf("%.17g\n", (double)post[1][i]);
        }
        printf("seconds_per_eval %.9g\n", std::chrono::duration<double>(t1 - t0).count() / repeat);
    }
    return 0;
}
