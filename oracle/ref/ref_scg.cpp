// ref_scg -- TEST INFRASTRUCTURE: runs the UNMODIFIED reference optimisers
// (medgpc/src/util/c_optimizer_scg.cpp, c_optimizer_varEM.cpp) on the analytic objective of
// tests/cpp/analytic_objective.h and on the reference's random initialisation, to generate the
// golden vectors under tests/golden/ (tests/golden/make_golden.py).
//   ref_scg scg   <max_iteration> <x0...>              -> loss, parameters
//   ref_scg varem <max_iteration> <sub_iter> Q D R eta beta_lam <x0 (D + Q(DR+2+D))...>
//   ref_scg init  <exp_setup.json> <count>              -> first <count> random theta vectors
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "core/gp_model_include.h"
#include "dataio/c_experiment.h"
#include "util/c_optimizer_scg.h"
#include "util/c_optimizer_varEM.h"
#include "../../tests/cpp/analytic_objective.h"

struct analytic_obj : public c_objective {
    int calls = 0;
    bool compute_objective(const bool &, const vector<double> &x, double &f, vector<double> &g, c_kernel *&,
                           c_meanfunc *&, c_likelihood *&, c_inference *&, c_prior *&) {
        calls++;
        double ff; vector<double> gg;
        if (!analytic_objective(x, ff, gg)) return false;
        f = ff; g = gg;
        return true;
    }
};

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    c_kernel *k = NULL; c_meanfunc *m = NULL; c_likelihood *l = NULL; c_inference *inf = NULL; c_prior *p = NULL;
    if (!strcmp(argv[1], "scg")) {
        const int iters = atoi(argv[2]);
        vector<double> x0;
        for (int i = 3; i < argc; i++) x0.push_back(atof(argv[i]));
        analytic_obj obj; c_optimizer_scg opt; double loss = 0; vector<double> out;
        opt.optimize(iters, x0, &obj, false, loss, out, k, m, l, inf, p);
        printf("calls %d\nloss %.17g\n", obj.calls, loss);
        for (size_t i = 0; i < out.size(); i++) printf("x %.17g\n", out[i]);
    } else if (!strcmp(argv[1], "varem")) {
        const int iters = atoi(argv[2]), sub = atoi(argv[3]);
        vector<int> kp = {atoi(argv[4]), atoi(argv[5]), atoi(argv[6])};
        vector<float> ph = {(float)atof(argv[7]), (float)atof(argv[8])};
        vector<double> x0;
        for (int i = 9; i < argc; i++) x0.push_back(atof(argv[i]));
        const int Q = kp[0], D = kp[1], R = kp[2], ncov = Q * (D * R + 2 + D);
        c_kernel_LMC_SM kernel(kp); c_meanfunc_zero mean; vector<int> lp = {D}; c_likelihood_gaussianMO lik(lp);
        c_prior prior(ncov, 0, D);
        prior.setup_param(7, kp, 2, ph);
        k = &kernel; m = &mean; l = &lik; p = &prior;
        analytic_obj obj; c_optimizer_varEM opt; opt.set_sub_opt_iter(sub); double loss = 0; vector<double> out;
        opt.optimize(iters, x0, &obj, false, loss, out, k, m, l, inf, p);
        printf("calls %d\nloss %.17g\n", obj.calls, loss);
        for (size_t i = 0; i < out.size(); i++) printf("x %.17g\n", out[i]);
        vector<double> ve = prior.get_cov_varEM_all();
        for (size_t i = 0; i < ve.size(); i++) printf("v %.17g\n", ve[i]);
        for (int i = 0; i < ncov; i++) printf("t %d\n", prior.type_cov[i]);
    } else if (!strcmp(argv[1], "init")) {
        c_experiment e(argv[2]);
        vector<vector<double> > hyp;
        e.get_global_hyp(hyp);
        const int cnt = atoi(argv[3]);
        for (int r = 0; r < cnt && r < (int)hyp.size(); r++)
            for (size_t i = 0; i < hyp[r].size(); i++) printf("h %.17g\n", hyp[r][i]);
    }
    return 0;
}
