// c_inference_cuda.cpp -- see c_inference_cuda.h.  Reference-side binding of libmedgp_cuda.so.
#include "c_inference_cuda.h"

#include <cmath>
#include <cstdlib>
#include <iostream>

#include "util/global_settings.h"  // PI = 3.14159265

using namespace std;

c_inference_cuda::c_inference_cuda() : ctx_(NULL), Q_(0), D_(0), R_(0), sid_(-1), status_(0), order_(-1)
{
    inffunc_name = "c_inference_cuda";
    inf_thread_num = 1;
}

c_inference_cuda::c_inference_cuda(const int &thread_num) : ctx_(NULL), Q_(0), D_(0), R_(0), sid_(-1), status_(0), order_(-1)
{
    inffunc_name = "c_inference_cuda";
    inf_thread_num = thread_num;  // kept for the CLI; the GPU path does not use host threads
}

c_inference_cuda::~c_inference_cuda()
{
    if (ctx_) medgp_cuda_destroy(ctx_);
}

static void die(const char *what, medgp_ctx *ctx, int rc)
{
    cout << "ERROR: " << what << " failed with status " << rc;
    if (ctx) cout << " (" << medgp_cuda_last_error(ctx) << ")";
    cout << "; libmedgp_cuda.so has no CPU fallback" << endl;
    exit(1);
}

// context + model shape; the kernel must be the SM-LMC kernel (kernel_param = Q, D, R)
void c_inference_cuda::bind(c_kernel *kernel)
{
    vector<int> kp = kernel->get_kernel_param();
    if (kp.size() != 3 || kernel->get_kernel_hyp_num() != kp[0] * (kp[1] * kp[2] + 2 + kp[1])) {
        cout << "ERROR: c_inference_cuda implements the LMC-SM kernel (kernel_index 7) only" << endl;
        exit(1);
    }
    if (!ctx_) {
        size_t workspace = (size_t)4 << 30;
        if (getenv("MEDGP_WORKSPACE_MB")) workspace = (size_t)atoll(getenv("MEDGP_WORKSPACE_MB")) << 20;
        int rc = medgp_cuda_create(&ctx_, 0, workspace);
        if (rc != MEDGP_OK) die("medgp_cuda_create", NULL, rc);
    }
    if (kp[0] != Q_ || kp[1] != D_ || kp[2] != R_) {
        medgp_cuda_clear_series(ctx_);
        sid_ = -1;
        int rc = medgp_cuda_model(ctx_, kp[0], kp[1], kp[2], PI);
        if (rc != MEDGP_OK) die("medgp_cuda_model", ctx_, rc);
        Q_ = kp[0]; D_ = kp[1]; R_ = kp[2];
    }
}

// the data set is uploaded once and reused for as long as the caller passes the same arrays
// (every evaluation of an optimiser run does)
void c_inference_cuda::upload(const vector<int> &meta, const vector<float> &x, const vector<float> &y, int order)
{
    if (sid_ >= 0 && order == order_ && meta == meta_ && x == x_ && y == y_) return;
    if (sid_ >= 0) medgp_cuda_free_series(ctx_, sid_);
    vector<int32_t> m(meta.begin(), meta.end());
    int rc = medgp_cuda_add_series_ordered(ctx_, (int)x.size(), &m[0], &x[0], &y[0], order, &sid_);
    if (rc != MEDGP_OK) die("medgp_cuda_add_series_ordered", ctx_, rc);
    meta_ = meta; x_ = x; y_ = y; order_ = order;
}

bool c_inference_cuda::compute_nlml(const bool &flag_grad, const vector<int> &meta, const vector<float> &x,
                                    const vector<float> &y, c_kernel *kernel, c_meanfunc *meanfunc,
                                    c_likelihood *likfunc, c_prior *prior, float *&chol_alpha,
                                    float *&chol_factor_inv, float &beta, double &nlml, vector<double> &dnlml)
{
    if (meanfunc->get_meanfunc_hyp_num() != 0) {
        cout << "ERROR: c_inference_cuda implements the zero mean function only" << endl;
        exit(1);
    }
    bind(kernel);
    // without gradient the caller wants the factor back, in ITS point order; with gradient the
    // library's feature-major order serves the fused gradient kernel
    upload(meta, x, y, flag_grad ? MEDGP_ORDER_FEATURE : MEDGP_ORDER_GIVEN);
    // theta = [log sigma (D) | A raw | log mu | log v | log kappa]: set_kernel_hyp / set_likfunc_hyp
    // keep the TRANSFORMED values only (c_kernel_LMC_SM.cpp:57-59, c_likelihood.cpp:41), so the
    // stored ones are recovered with log()
    vector<double> lik_hyp = likfunc->get_likfunc_hyp(), cov_hyp = kernel->get_kernel_hyp();
    theta_.clear();
    for (size_t i = 0; i < lik_hyp.size(); i++) theta_.push_back(log(lik_hyp[i]));
    for (size_t i = 0; i < cov_hyp.size(); i++) theta_.push_back((int)i < Q_ * D_ * R_ ? cov_hyp[i] : log(cov_hyp[i]));
    if ((int)theta_.size() != medgp_cuda_num_hyp(ctx_)) {
        cout << "ERROR: mismatch # of hyperparameters! Get " << theta_.size() << ", but expect "
             << medgp_cuda_num_hyp(ctx_) << endl;
        exit(1);
    }
    vector<double> grad(theta_.size(), 0.0);
    double value = 0.0;
    if (flag_grad) {
        int rc = medgp_cuda_nlml_grad(ctx_, 1, &sid_, &theta_[0], 1, &value, &grad[0], &status_);
        if (rc != MEDGP_OK) die("medgp_cuda_nlml_grad", ctx_, rc);
    } else {
        // chol_alpha / chol_factor_inv are the caller's n*n float buffers (gp_regression.cpp:116-117)
        int rc = medgp_cuda_export_factors(ctx_, sid_, &theta_[0], chol_alpha, chol_factor_inv, &value, &status_);
        if (rc != MEDGP_OK) die("medgp_cuda_export_factors", ctx_, rc);
        if (status_ >= 0) {
            double quad = 0.0;
            for (size_t i = 0; i < y.size(); i++) quad += (double)y[i] * (double)chol_alpha[i];
            beta = (float)quad;  // c_inference_exact.cpp:146-147
        }
    }
    if (status_ > 0) cout << "WARNING: Cholesky decomposition failed! jittered " << status_ << " time(s)" << endl;
    if (status_ < 0) return false;  // as spotrf failing after 10 additions (c_inference_exact.cpp:109-111)
    nlml = value;
    dnlml.clear();
    if (flag_grad) dnlml = grad;

    // prior terms, as inference/c_inference_prior.cpp:59-150: nlml -= log p(h); the gradient is
    // w.r.t. the stored value, so exp-transformed hyper-parameters pick up the factor h
    if (prior != NULL) {
        int offset = 0;
        for (int i = 0; i < (int)lik_hyp.size(); i++) {
            if (!prior->flag_lik[i]) continue;
            if (prior->type_lik[i] == 0) {
                if (flag_grad) dnlml[i + offset] = 0.0;
                continue;
            }
            vector<double> lp = prior->get_one_lik_lik(lik_hyp[i], i);
            nlml -= lp[0];
            if (flag_grad) dnlml[i + offset] -= prior->exp_lik[i] ? lik_hyp[i] * lp[1] : lp[1];
        }
        offset += (int)lik_hyp.size();
        for (int i = 0; i < (int)cov_hyp.size(); i++) {
            if (!prior->flag_cov[i]) continue;
            if (prior->type_cov[i] == 0) {
                if (flag_grad) dnlml[i + offset] = 0.0;
                continue;
            }
            if (prior->type_cov[i] == -1) continue;
            vector<double> lp = prior->get_one_lik_cov(cov_hyp[i], i);
            nlml -= lp[0];
            if (flag_grad) dnlml[i + offset] -= prior->exp_cov[i] ? cov_hyp[i] * lp[1] : lp[1];
        }
        // zero mean: no mean hyper-parameters
    }
    return true;
}

bool c_inference_cuda::predict(const vector<int> &meta2, const vector<float> &x2, vector<double> &mean,
                               vector<double> &var)
{
    if (!ctx_ || sid_ < 0 || theta_.empty()) return false;
    const int m = (int)x2.size();
    mean.assign(m, 0.0);
    var.assign(m, 0.0);
    if (m == 0) return true;
    vector<int32_t> ms(meta2.begin(), meta2.end());
    const int off[2] = {0, m};
    int status = 0;
    int rc = medgp_cuda_predict(ctx_, 1, &sid_, &theta_[0], off, &ms[0], &x2[0], &mean[0], &var[0], &status);
    if (rc != MEDGP_OK) die("medgp_cuda_predict", ctx_, rc);
    return status >= 0;
}
