// host_check -- test driver for OUR host classes (medgp_b200/host), same commands and output
// format as oracle/ref/ref_scg.cpp so tests can diff the two.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
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
    } else if (!strcmp(argv[1], "dscg") || !strcmp(argv[1], "dvarem")) {
        // the DEVICE-resident optimiser (medgp_cuda_scg_*: state machine in HBM) on the analytic
        // objective, driven through the session's taps by medgp_optimize_on_device; same output
        // format as "scg" / "varem".  The model only fixes P: unused trailing coordinates get a
        // zero gradient, which leaves every dot product -- and so the trajectory -- unchanged.
        const bool vem = !strcmp(argv[1], "dvarem");
        const int iters = atoi(argv[2]);
        int sub = 0, a0 = 3;
        vector<int> kp = {1, 2, 1};
        vector<float> ph = {0.01f, 0.01f};
        if (vem) {
            sub = atoi(argv[3]);
            kp = {atoi(argv[4]), atoi(argv[5]), atoi(argv[6])};
            ph = {(float)atof(argv[7]), (float)atof(argv[8])};
            a0 = 9;
        }
        vector<double> x0;
        for (int i = a0; i < argc; i++) x0.push_back(atof(argv[i]));
        const int Q = kp[0], D = kp[1], R = kp[2], ncov = Q * (D * R + 2 + D);
        medgp_ctx *ctx = medgp_backend::context(Q, D, R);
        const int P = medgp_cuda_num_hyp(ctx);
        if ((int)x0.size() > P) { printf("error: x0 longer than P\n"); return 2; }
        struct user_t { size_t n; int calls; } user = {x0.size(), 0};
        c_prior prior(ncov, 0, D);
        if (vem) prior.setup_param(7, kp, 2, ph);
        // two identical instances: the batch must advance them independently and identically
        std::vector<medgp_opt_instance> inst(2);
        for (auto &it : inst) {
            it.series_id = 0;
            it.init_parameter = x0;
            it.init_parameter.resize(P, 0.0);
            it.prior = vem ? &prior : nullptr;
            it.use_varem = vem;
            it.max_iteration = iters;
            it.sub_opt_iter = sub;
        }
        c_prior prior2 = prior;
        if (vem) inst[1].prior = &prior2;
        // a dummy series so that series_id 0 exists
        { int sid; int32_t m3[3] = {0, 0, 1 % D}; float x3[3] = {1, 2, 3}, y3[3] = {0.1f, 0.2f, 0.3f};
          medgp_cuda_add_series(ctx, 3, m3, x3, y3, &sid); for (auto &it : inst) it.series_id = sid; }
        medgp_optimize_on_device(ctx, kp, D, inst, 16,
            [](int index, const vector<double> &x, double &f, vector<double> &g, void *u) -> bool {
                user_t *us = (user_t *)u;
                if (index == 0) us->calls++;
                vector<double> xs(x.begin(), x.begin() + us->n), gs;
                if (!analytic_objective(xs, f, gs)) return false;
                g.assign(x.size(), 0.0);
                std::copy(gs.begin(), gs.end(), g.begin());
                return true;
            }, &user);
        for (int k = 0; k < P; k++)
            if (inst[0].opt_parameter[k] != inst[1].opt_parameter[k]) { printf("error: instances differ\n"); return 3; }
        printf("calls %d\nloss %.17g\n", user.calls, inst[0].opt_loss);
        for (size_t i = 0; i < x0.size(); i++) printf("x %.17g\n", inst[0].opt_parameter[i]);
        if (vem) {
            vector<double> ve = prior.get_cov_varEM_all();
            for (size_t i = 0; i < ve.size(); i++) printf("v %.17g\n", ve[i]);
            for (int i = 0; i < ncov; i++) printf("t %d\n", prior.type_cov[i]);
        }
        medgp_backend::shutdown();
    } else if (!strcmp(argv[1], "lpt")) {
        // lpt <nshard> <size>...: the C++ deal of patients to shards (front-ends)
        vector<int> sizes;
        for (int i = 3; i < argc; i++) sizes.push_back(atoi(argv[i]));
        const vector<int> sh = medgp_lpt_assign(sizes, atoi(argv[2]));
        for (size_t i = 0; i < sh.size(); i++) printf("s %d\n", sh[i]);
    } else if (!strcmp(argv[1], "sizes")) {
        // sizes <cfg> <pan>...: sizes-only pass and full cohort load of c_experiment
        c_experiment e(argv[2]);
        vector<std::string> pans;
        for (int i = 3; i < argc; i++) pans.push_back(argv[i]);
        const vector<int> sz = e.get_cohort_sizes(pans);
        vector<c_experiment::patient_data> data;
        e.get_cohort_data(pans, data);
        for (size_t k = 0; k < pans.size(); k++) {
            vector<int> m; vector<float> t, v;
            e.get_one_patient_data(pans[k], m, t, v, false);
            const bool same = m == data[k].meta && t == data[k].time && v == data[k].value && (int)t.size() == sz[k];
            printf("n %d\nsame %d\n", sz[k], (int)same);
        }
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
