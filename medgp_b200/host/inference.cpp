// inference.cpp -- the drop-in seam: c_inference_exact / c_inference_prior / GP_Regression /
// c_objective_one forward to libmedgp_cuda.so (include/medgp_cuda.h).  See medgp_host.h.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <list>
#include <map>

#include "medgp_host.h"

using std::vector;

// ------------------------------------------------------------------ backend
namespace {
struct BackendSlot {
    medgp_ctx *ctx = nullptr;
    int Q = 0, D = 0, R = 0;
    // small LRU of uploaded series keyed by a content hash
    struct Entry {
        uint64_t key; int id;
        std::vector<int> meta;       // the uploaded arrays themselves: a hash hit is confirmed by content
        std::vector<float> x, y;
    };
    std::list<Entry> lru;
};
std::map<int, BackendSlot> g_slots;
const size_t kSeriesCache = 64;

uint64_t fnv(uint64_t h, const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
    return h;
}

BackendSlot *slot_of(medgp_ctx *ctx)
{
    for (auto &kv : g_slots)
        if (kv.second.ctx == ctx) return &kv.second;
    return nullptr;
}

void die(const char *what, medgp_ctx *ctx, int rc)
{
    std::cerr << "ERROR: " << what << " failed with status " << rc;
    if (ctx) std::cerr << " (" << medgp_cuda_last_error(ctx) << ")";
    std::cerr << "; libmedgp_cuda.so has no CPU fallback" << std::endl;
    exit(1);
}
}  // namespace

medgp_ctx *medgp_backend::context(int Q, int D, int R, int device)
{
    BackendSlot &s = g_slots[device];
    if (!s.ctx) {
        size_t ws = 0;
        if (const char *e = getenv("MEDGP_WORKSPACE_MB")) ws = (size_t)atoll(e) << 20;
        else ws = (size_t)4096 << 20;
        int rc = medgp_cuda_create(&s.ctx, device, ws);
        if (rc != MEDGP_OK) die("medgp_cuda_create", nullptr, rc);
    }
    if (s.Q != Q || s.D != D || s.R != R) {
        drop_series_cache(s.ctx);
        medgp_cuda_clear_series(s.ctx);
        int rc = medgp_cuda_model(s.ctx, Q, D, R, PI);
        if (rc != MEDGP_OK) die("medgp_cuda_model", s.ctx, rc);
        s.Q = Q; s.D = D; s.R = R;
    }
    return s.ctx;
}

void medgp_backend::shutdown()
{
    for (auto &kv : g_slots)
        if (kv.second.ctx) medgp_cuda_destroy(kv.second.ctx);
    g_slots.clear();
}

int medgp_backend::series_for(medgp_ctx *ctx, const vector<int> &meta, const vector<float> &x,
                              const vector<float> &y)
{
    BackendSlot *s = slot_of(ctx);
    uint64_t key = 1469598103934665603ULL;
    key = fnv(key, meta.data(), meta.size() * sizeof(int));
    key = fnv(key, x.data(), x.size() * sizeof(float));
    key = fnv(key, y.data(), y.size() * sizeof(float));
    if (s)
        for (auto it = s->lru.begin(); it != s->lru.end(); ++it)
            if (it->key == key && it->meta == meta && it->x == x && it->y == y) {
                s->lru.splice(s->lru.begin(), s->lru, it);
                return s->lru.front().id;
            }
    int id = -1;
    static_assert(sizeof(int) == sizeof(int32_t), "meta is int32");
    int rc = medgp_cuda_add_series(ctx, (int)x.size(), (const int32_t *)meta.data(), x.data(), y.data(), &id);
    if (rc != MEDGP_OK) die("medgp_cuda_add_series", ctx, rc);
    if (s) {
        s->lru.push_front({key, id, meta, x, y});
        if (s->lru.size() > kSeriesCache) {
            medgp_cuda_free_series(ctx, s->lru.back().id);
            s->lru.pop_back();
        }
    }
    return id;
}

void medgp_backend::drop_series_cache(medgp_ctx *ctx)
{
    BackendSlot *s = slot_of(ctx);
    if (s) s->lru.clear();
}

// ------------------------------------------------------------------ inference
void c_inference::print_inffunc() const
{
    std::cout << "current inference object: " << inffunc_name << "; number of threads: " << inf_thread_num
              << " (GPU backend: libmedgp_cuda.so)" << std::endl;
}

bool c_inference_exact::compute_nlml(const bool &flag_grad, const vector<int> &meta, const vector<float> &x,
                                     const vector<float> &y, c_kernel *kernel, c_meanfunc *meanfunc,
                                     c_likelihood *likfunc, c_prior *, medgp_fit &fit, double &nlml,
                                     vector<double> &dnlml)
{
    c_kernel_LMC_SM *lmc = dynamic_cast<c_kernel_LMC_SM *>(kernel);
    if (!lmc) {
        std::cout << "ERROR: the GPU backend implements the LMC-SM kernel (kernel_index 7) only" << std::endl;
        exit(1);
    }
    if (meanfunc->get_meanfunc_hyp_num() != 0) {
        std::cout << "ERROR: the GPU backend implements the zero mean function only" << std::endl;
        exit(1);
    }
    const vector<int> kp = kernel->get_kernel_param();
    medgp_ctx *ctx = medgp_backend::context(kp[0], kp[1], kp[2]);
    const int sid = medgp_backend::series_for(ctx, meta, x, y);
    vector<double> theta = likfunc->get_likfunc_hyp_raw();
    const vector<double> cov = kernel->get_kernel_hyp_raw();
    theta.insert(theta.end(), cov.begin(), cov.end());
    if ((int)theta.size() != medgp_cuda_num_hyp(ctx)) {
        std::cout << "ERROR: mismatch # of hyperparameters! Get " << theta.size() << ", but expect "
                  << medgp_cuda_num_hyp(ctx) << std::endl;
        exit(1);
    }
    vector<double> grad(flag_grad ? theta.size() : 0);
    int status = 0;
    double value = 0.0;
    int rc = medgp_cuda_nlml_grad(ctx, 1, &sid, theta.data(), flag_grad ? 1 : 0, &value,
                                  flag_grad ? grad.data() : nullptr, &status);
    if (rc != MEDGP_OK) die("medgp_cuda_nlml_grad", ctx, rc);
    if (status > 0)
        std::cout << "WARNING: Cholesky decomposition failed! jittered " << status << " time(s)" << std::endl;
    if (status < 0) return false;  // still not positive definite after 10 jitters
    nlml = value;
    if (flag_grad) dnlml = grad; else dnlml.clear();
    fit.ctx = ctx;
    fit.series_id = sid;
    fit.theta = theta;
    fit.status = status;
    return true;
}

bool c_inference_prior::compute_nlml(const bool &flag_grad, const vector<int> &meta, const vector<float> &x,
                                     const vector<float> &y, c_kernel *kernel, c_meanfunc *meanfunc,
                                     c_likelihood *likfunc, c_prior *prior, medgp_fit &fit, double &nlml,
                                     vector<double> &dnlml)
{
    c_inference_exact major_inffunc(inf_thread_num);
    const bool ok = major_inffunc.compute_nlml(flag_grad, meta, x, y, kernel, meanfunc, likfunc, prior, fit,
                                               nlml, dnlml);
    if (ok && prior != NULL)
        medgp_apply_prior(*prior, likfunc->get_likfunc_hyp(), kernel->get_kernel_hyp(),
                          meanfunc->get_meanfunc_hyp(), flag_grad, nlml, dnlml);
    return ok;
}

// ------------------------------------------------------------------ GP_Regression
GP_Regression::GP_Regression()
    : dim(0), flag_trained(false), nlm_likelihood(0.0), kernel(nullptr), meanfunc(nullptr),
      likfunc(nullptr), inffunc(nullptr), prior(nullptr)
{
}

GP_Regression::GP_Regression(const int &input_dim, c_kernel *k, c_meanfunc *m, c_likelihood *l,
                             c_inference *inf, c_prior *p)
{
    reset(input_dim, k, m, l, inf, p);
}

void GP_Regression::reset(const int &input_dim, c_kernel *k, c_meanfunc *m, c_likelihood *l,
                          c_inference *inf, c_prior *p)
{
    dim = input_dim;
    flag_trained = false;
    kernel = k; meanfunc = m; likfunc = l; inffunc = inf; prior = p;
    nlm_likelihood = 0.0;
    dnlm_likelihood.clear();
    fit = medgp_fit();
}

void GP_Regression::train(const bool &flag_grad, const vector<int> &meta, const vector<float> &x,
                          const vector<float> &y)
{
    flag_trained = inffunc->compute_nlml(flag_grad, meta, x, y, kernel, meanfunc, likfunc, prior, fit,
                                         nlm_likelihood, dnlm_likelihood);
    if (!flag_trained) std::cout << "Warning: current inference failed in train()!!" << std::endl;
}

vector<vector<float> > GP_Regression::predict(const vector<int> &meta, const vector<int> &meta2,
                                              const vector<float> &x, const vector<float> &y,
                                              const vector<float> &x2)
{
    if (!flag_trained) train(false, meta, x, y);
    const int m = (int)x2.size();
    vector<vector<float> > posterior(2, vector<float>(m, 0.0f));
    if (!flag_trained || m == 0) return posterior;
    vector<double> mean(m), var(m);
    const int off[2] = {0, m};
    int status = 0;
    // the series cache may have evicted (and reused) the id since train(): resolve it again from
    // the arrays, which uploads them anew if they are gone
    fit.series_id = medgp_backend::series_for(fit.ctx, meta, x, y);
    int rc = medgp_cuda_predict(fit.ctx, 1, &fit.series_id, fit.theta.data(), off,
                                (const int32_t *)meta2.data(), x2.data(), mean.data(), var.data(), &status);
    if (rc != MEDGP_OK) die("medgp_cuda_predict", fit.ctx, rc);
    for (int i = 0; i < m; i++) {
        posterior[0][i] = (float)mean[i];
        posterior[1][i] = (float)var[i];
    }
    return posterior;
}

// ------------------------------------------------------------------ objectives
void c_objective::print_objective() const { std::cout << "current objective: " << objective_name << std::endl; }

c_objective_one::c_objective_one() : obj_kernel_idx(7) { objective_name = "c_objective_one"; }

c_objective_one::c_objective_one(const int &kernel_idx, const vector<int> &kernel_param,
                                 const vector<int> &meta, const vector<float> &x, const vector<float> &y)
    : obj_meta(meta), obj_x(x), obj_y(y), obj_kernel_idx(kernel_idx), obj_kernel_param(kernel_param)
{
    objective_name = "c_objective_one";
}

bool c_objective_one::compute_objective(const bool &flag_grad, const vector<double> &input_parameter,
                                        double &objective_value, vector<double> &gradients,
                                        c_kernel *&input_kernel, c_meanfunc *&input_meanfunc,
                                        c_likelihood *&input_likfunc, c_inference *&input_inffunc,
                                        c_prior *&input_prior)
{
    if ((int)obj_x.size() <= 2) return false;  // util/c_objective_one.cpp:51
    c_hyperparam hyp(input_parameter, input_kernel->get_kernel_hyp_num(),
                     input_meanfunc->get_meanfunc_hyp_num(), input_likfunc->get_likfunc_hyp_num());
    input_kernel->set_kernel_hyp(hyp.get_hyp_cov());
    input_meanfunc->set_meanfunc_hyp(hyp.get_hyp_mean());
    input_likfunc->set_likfunc_hyp(hyp.get_hyp_lik());
    GP_Regression model(1, input_kernel, input_meanfunc, input_likfunc, input_inffunc, input_prior);
    model.train(flag_grad, obj_meta, obj_x, obj_y);
    if (!model.get_flag_trained()) return false;
    objective_value = model.get_neg_log_mlikelihood();
    if (flag_grad) gradients = model.get_dneg_log_mlikelihood();
    return true;
}

void c_objective_batch::release()
{
    if (!cap_) return;
    medgp_cuda_host_free(ctx_, h_theta_);
    medgp_cuda_host_free(ctx_, h_nlml_);
    medgp_cuda_host_free(ctx_, h_grad_);
    medgp_cuda_host_free(ctx_, h_status_);
    h_theta_ = h_nlml_ = h_grad_ = nullptr;
    h_status_ = nullptr;
    cap_ = 0;
}

void c_objective_batch::compute(const bool &flag_grad, const vector<medgp_eval_request> &reqs,
                                vector<medgp_eval_result> &out)
{
    const int B = (int)reqs.size();
    const int P = medgp_cuda_num_hyp(ctx_);
    out.assign(B, medgp_eval_result());
    if (B == 0) return;
    if ((size_t)B > cap_) {
        release();
        cap_ = std::max<size_t>(2 * (size_t)B, 64);
        void *p[4] = {nullptr, nullptr, nullptr, nullptr};
        const size_t bytes[4] = {cap_ * P * sizeof(double), cap_ * sizeof(double), cap_ * P * sizeof(double), cap_ * sizeof(int)};
        for (int i = 0; i < 4; i++) {
            const int rc = medgp_cuda_host_alloc(ctx_, bytes[i], &p[i]);
            if (rc != MEDGP_OK) die("medgp_cuda_host_alloc", ctx_, rc);
        }
        h_theta_ = (double *)p[0]; h_nlml_ = (double *)p[1]; h_grad_ = (double *)p[2]; h_status_ = (int *)p[3];
    }
    vector<int> sids(B);
    double *theta = h_theta_, *nlml = h_nlml_, *grad = h_grad_;
    int *status = h_status_;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        sids[b] = reqs[b].series_id;
        memcpy(&theta[(size_t)b * P], reqs[b].theta->data(), sizeof(double) * P);
    }
    int rc = medgp_cuda_nlml_grad(ctx_, B, sids.data(), theta, flag_grad ? 1 : 0, nlml, flag_grad ? grad : nullptr, status);
    if (rc != MEDGP_OK) die("medgp_cuda_nlml_grad", ctx_, rc);
    const int nlik = D_, nA = Q_ * D_ * R_;
    // per-request epilogue (copy out, prior terms): independent requests, read-only priors
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        medgp_eval_result &r = out[b];
        r.status = status[b];
        r.ok = status[b] >= 0;
        if (!r.ok) continue;
        r.value = nlml[b];
        if (flag_grad) r.grad.assign(grad + (size_t)b * P, grad + (size_t)(b + 1) * P);
        if (reqs[b].prior) {
            const double *th = &theta[(size_t)b * P];
            vector<double> lik(nlik), cov(P - nlik), mean;
            for (int i = 0; i < nlik; i++) lik[i] = exp(th[i]);
            for (int i = 0; i < P - nlik; i++) cov[i] = i < nA ? th[nlik + i] : exp(th[nlik + i]);
            medgp_apply_prior(*reqs[b].prior, lik, cov, mean, flag_grad, r.value, r.grad);
        }
    }
}
