// oracle_backend_scg.cpp -- TEST INFRASTRUCTURE ONLY (see oracle_backend.c).
//
// The optimiser-session entry points of include/medgp_cuda.h (medgp_cuda_scg_*) on top of the
// CPU oracle and the HOST re-entrant SCG stepper (medgp_b200/host/optimizer.cpp), so that the
// host-side driver of the device-resident optimiser (medgp_optimize_on_device: variational-EM
// rounds, prior tables, result handling) can be exercised in the GPU-less build container.
// The device state machine itself (medgp_b200/csrc/scg.cuh) is tested on the GPU against the
// reference's trajectories (tests/test_gpu_optimizer.py).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../medgp_b200/host/medgp_host.h"

struct medgp_scg {
    medgp_ctx *ctx;
    int count, P, n_lik, n_a;
    std::vector<int> series;
    std::vector<scg_stepper> st;
    std::vector<signed char> ptype, pexp;
    std::vector<float> ppar;
    bool has_prior;
};

extern "C" {

int medgp_cuda_scg_create(medgp_ctx *ctx, int count, medgp_scg **out)
{
    if (!ctx || count < 1 || !out) return MEDGP_ERR_ARG;
    medgp_scg *g = new medgp_scg();
    g->ctx = ctx;
    g->count = count;
    g->P = medgp_cuda_num_hyp(ctx);
    g->has_prior = false;
    *out = g;
    return MEDGP_OK;
}

void medgp_cuda_scg_destroy(medgp_scg *g) { delete g; }

int medgp_cuda_scg_start(medgp_scg *g, const int *series_id, const double *theta0, const int *max_iteration,
                         const signed char *prior_type, const signed char *prior_exp, const float *prior_param)
{
    const size_t P = g->P;
    g->series.assign(series_id, series_id + g->count);
    g->st.clear();
    for (int b = 0; b < g->count; b++)
        g->st.push_back(scg_stepper(max_iteration[b], std::vector<double>(theta0 + b * P, theta0 + (b + 1) * P)));
    g->has_prior = prior_type != nullptr;
    if (g->has_prior) {
        g->ptype.assign(prior_type, prior_type + g->count * P);
        g->pexp.assign(prior_exp, prior_exp + g->count * P);
        g->ppar.assign(prior_param, prior_param + g->count * P * 2);
    }
    return MEDGP_OK;
}

// prior terms from the flat table, same arithmetic as medgp_apply_prior / c_prior (host)
static void apply_table(const medgp_scg *g, int b, int n_lik, int n_a, const double *theta, double &f, std::vector<double> &grad)
{
    const size_t P = g->P;
    for (size_t k = 0; k < P; k++) {
        const int t = g->ptype[b * P + k];
        if (t < 0) continue;
        if (t == 0) { grad[k] = 0.0; continue; }
        const bool raw = (int)k >= n_lik && (int)k < n_lik + n_a;
        const double h = raw ? theta[k] : exp(theta[k]);
        const std::vector<float> par = {g->ppar[2 * (b * P + k)], g->ppar[2 * (b * P + k) + 1]};
        const std::vector<double> lp = t == 1 ? c_prior::prior_lik_normal(h, par) : c_prior::prior_lik_laplace(h, par);
        f -= lp[0];
        grad[k] -= g->pexp[b * P + k] ? h * lp[1] : lp[1];
    }
}

// model shape of the oracle backend's context (oracle_backend.c keeps Q, D, R, P first)
struct ctx_head { int Q, D, R, P; };

int medgp_cuda_scg_run(medgp_scg *g, int super_steps, int *active_left)
{
    const ctx_head *h = (const ctx_head *)g->ctx;
    const size_t P = g->P;
    for (int step = 0; step < super_steps; step++) {
        std::vector<int> who;
        for (int b = 0; b < g->count; b++)
            if (g->st[b].wants_eval()) who.push_back(b);
        if (who.empty()) break;
        const int B = (int)who.size();
        std::vector<int> sids(B), status(B);
        std::vector<double> theta(B * P), f(B), grad(B * P);
        for (int q = 0; q < B; q++) {
            sids[q] = g->series[who[q]];
            memcpy(&theta[q * P], g->st[who[q]].point().data(), P * sizeof(double));
        }
        int rc = medgp_cuda_nlml_grad(g->ctx, B, sids.data(), theta.data(), 1, f.data(), grad.data(), status.data());
        if (rc) return rc;
        for (int q = 0; q < B; q++) {
            std::vector<double> gq(grad.begin() + q * P, grad.begin() + (q + 1) * P);
            double fq = f[q];
            const bool ok = status[q] >= 0;
            if (ok && g->has_prior) apply_table(g, who[q], h->D, h->Q * h->D * h->R, &theta[q * P], fq, gq);
            g->st[who[q]].feed(ok, fq, gq);
        }
    }
    int left = 0;
    for (int b = 0; b < g->count; b++) left += g->st[b].wants_eval() ? 1 : 0;
    if (active_left) *active_left = left;
    return MEDGP_OK;
}

int medgp_cuda_scg_result(medgp_scg *g, double *theta_best, double *loss, int *evals)
{
    const size_t P = g->P;
    for (int b = 0; b < g->count; b++) {
        if (theta_best) memcpy(theta_best + b * P, g->st[b].best_parameter().data(), P * sizeof(double));
        if (loss) loss[b] = g->st[b].best_loss();
        if (evals) evals[b] = g->st[b].evaluations();
    }
    return MEDGP_OK;
}

int medgp_cuda_scg_points(medgp_scg *g, double *theta, int *wants)
{
    const size_t P = g->P;
    for (int b = 0; b < g->count; b++) {
        if (theta) memcpy(theta + b * P, g->st[b].point().data(), P * sizeof(double));
        if (wants) wants[b] = g->st[b].wants_eval() ? 1 : 0;
    }
    return MEDGP_OK;
}

int medgp_cuda_scg_feed(medgp_scg *g, const double *f, const double *grad, const int *ok)
{
    const ctx_head *h = (const ctx_head *)g->ctx;
    const size_t P = g->P;
    for (int b = 0; b < g->count; b++) {
        if (!g->st[b].wants_eval()) continue;
        std::vector<double> gq(grad + b * P, grad + (b + 1) * P);
        double fq = f[b];
        if (ok[b] && g->has_prior) apply_table(g, b, h->D, h->Q * h->D * h->R, g->st[b].point().data(), fq, gq);
        g->st[b].feed(ok[b] != 0, fq, gq);
    }
    return MEDGP_OK;
}

}  // extern "C"
