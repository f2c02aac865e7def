// host_check -- test driver for OUR host classes (medgp_b200/host), same commands and output
// format as oracle/ref/ref_scg.cpp so tests can diff the two.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../medgp_b200/host/c_experiment.h"
#include "../../medgp_b200/host/medgp_host.h"
#include "analytic_objective.h"

using std::vector;

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
    } else if (!strcmp(argv[1], "prior")) {
        // prior <type> <x> <p0> <p1> : log density and derivative
        vector<float> par = {(float)atof(argv[4]), (float)atof(argv[5])};
        vector<double> r = atoi(argv[2]) == 1 ? c_prior::prior_lik_normal(atof(argv[3]), par)
                                              : c_prior::prior_lik_laplace(atof(argv[3]), par);
        printf("lp %.17g\ndlp %.17g\n", r[0], r[1]);
    }
    return 0;
}
