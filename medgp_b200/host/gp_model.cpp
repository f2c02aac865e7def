// gp_model.cpp -- hyper-parameter containers, SM-LMC kernel parameters, likelihood, priors.
// O(P) scalar work that the reference also keeps on the host; see medgp_host.h for the map to
// the reference files.
#include <cmath>
#include <cstdlib>
#include <iostream>

#include "medgp_host.h"

using std::vector;

// ------------------------------------------------------------------ c_hyperparam
c_hyperparam::c_hyperparam(const vector<double> &hyp_all, const int &num_cov, const int &num_mean,
                           const int &num_lik)
{
    set_hyp_all(hyp_all, num_cov, num_mean, num_lik);
}

c_hyperparam::c_hyperparam(const vector<double> &cov, const vector<double> &mean,
                           const vector<double> &lik)
    : hyp_cov(cov), hyp_mean(mean), hyp_lik(lik)
{
}

vector<double> c_hyperparam::get_hyp_all() const
{
    vector<double> all(hyp_lik);
    all.insert(all.end(), hyp_cov.begin(), hyp_cov.end());
    all.insert(all.end(), hyp_mean.begin(), hyp_mean.end());
    return all;
}

void c_hyperparam::set_hyp_all(const vector<double> &hyp, const int &num_cov, const int &num_mean,
                               const int &num_lik)
{
    // order [lik | cov | mean]
    hyp_lik.assign(hyp.begin(), hyp.begin() + num_lik);
    hyp_cov.assign(hyp.begin() + num_lik, hyp.begin() + num_lik + num_cov);
    hyp_mean.assign(hyp.begin() + num_lik + num_cov, hyp.begin() + num_lik + num_cov + num_mean);
}

// ------------------------------------------------------------------ kernels
void c_kernel::print_kernel() const
{
    std::cout << "current kernel object: " << kernel_name << "; # of hyperparameters: " << kernel_hyp_num
              << std::endl;
}

c_kernel_LMC_SM::c_kernel_LMC_SM() { kernel_name = "c_kernel_LMC_SM"; }

c_kernel_LMC_SM::c_kernel_LMC_SM(const vector<int> &input_param)
{
    kernel_name = "c_kernel_LMC_SM";
    set_kernel_param(input_param);
}

c_kernel_LMC_SM::c_kernel_LMC_SM(const vector<int> &input_param, const vector<double> &input_hyp)
{
    kernel_name = "c_kernel_LMC_SM";
    if (input_param.size() != 3) {
        std::cout << "ERROR:current input parameters should report 3 numbers (mixture, output, rank); "
                  << "received " << input_param.size() << std::endl;
        exit(1);
    }
    set_kernel_param(input_param);
    if ((int)input_hyp.size() != kernel_hyp_num) {
        std::cout << "ERROR: mismatch # of hyperparameters! Get " << input_hyp.size() << ", but expect "
                  << kernel_hyp_num << std::endl;
        exit(1);
    }
    set_kernel_hyp(input_hyp);
}

void c_kernel_LMC_SM::set_kernel_param(const vector<int> &input_param)
{
    kernel_param = input_param;
    const int Q = input_param[0], D = input_param[1], R = input_param[2];
    kernel_hyp_num = Q * (D * R + 2 + D);
}

void c_kernel_LMC_SM::set_kernel_hyp(const vector<double> &input_hyp)
{
    const int Q = kernel_param[0], D = kernel_param[1], R = kernel_param[2];
    kernel_hyp_raw = input_hyp;
    kernel_hyp = input_hyp;
    // everything after the A block is stored as a log
    for (size_t i = (size_t)Q * D * R; i < input_hyp.size(); i++) kernel_hyp[i] = exp(kernel_hyp[i]);
    compute_coregional_matrix();
}

void c_kernel_LMC_SM::compute_coregional_matrix()
{
    const int Q = kernel_param[0], D = kernel_param[1], R = kernel_param[2];
    coregional_matrix.assign(Q, vector<double>(D * D, 0.0));
    for (int q = 0; q < Q; q++) {
        const double *A = &kernel_hyp[q * D * R];
        const double *kappa = &kernel_hyp[Q * (D * R + 2) + q * D];
        vector<double> &B = coregional_matrix[q];
        for (int i = 0; i < D; i++)
            for (int j = 0; j < D; j++) {
                double s = 0.0;
                for (int r = 0; r < R; r++) s += A[i * R + r] * A[j * R + r];
                B[i * D + j] = s + (i == j ? kappa[i] : 0.0);
            }
    }
}

double c_kernel_LMC_SM::compute_k(const double &rsq, const double &mu, const double &v)
{
    return cos(2.0 * PI * sqrt(rsq) * mu) * exp(-2.0 * pow(PI * v, 2.0) * rsq);
}

double c_kernel_LMC_SM::compute_km(const double &rsq, const double &mu, const double &v)
{
    const double dmu = 2.0 * PI * sqrt(rsq) * mu;
    return -dmu * sin(dmu) * exp(-2.0 * pow(PI * v, 2.0) * rsq);
}

double c_kernel_LMC_SM::compute_kv(const double &rsq, const double &mu, const double &v)
{
    const double d2piv = pow(PI * v, 2.0) * rsq;
    return -4.0 * d2piv * cos(2.0 * PI * sqrt(rsq) * mu) * exp(-2.0 * d2piv);
}

// ------------------------------------------------------------------ likelihood / mean
void c_likelihood::print_likfunc() const
{
    std::cout << "current likelihood function object: " << likfunc_name << std::endl;
}

void c_likelihood::set_likfunc_hyp(vector<double> input_hyp)
{
    likfunc_hyp_raw = input_hyp;
    likfunc_hyp = input_hyp;
    for (size_t i = 0; i < input_hyp.size(); i++) likfunc_hyp[i] = exp(likfunc_hyp[i]);
}

c_likelihood_gaussianMO::c_likelihood_gaussianMO()
{
    likfunc_name = "c_likelihood_gaussianMO";
    likfunc_hyp_num = 1;
    std::cout << "WARNING: using multi-output likelihood function but no output number is specified! "
              << "Using default (1)" << std::endl;
}

c_likelihood_gaussianMO::c_likelihood_gaussianMO(vector<int> input_param) : c_likelihood(input_param)
{
    likfunc_name = "c_likelihood_gaussianMO";
    likfunc_hyp_num = input_param[0];
}

void c_meanfunc::print_meanfunc() const
{
    std::cout << "current mean function object: " << meanfunc_name << std::endl;
}

// ------------------------------------------------------------------ c_prior
c_prior::c_prior(int num_cov, int num_mean, int num_lik) { initialize_param(num_cov, num_mean, num_lik); }

void c_prior::initialize_param(int num_cov, int num_mean, int num_lik)
{
    hyp_cov_num = num_cov;
    hyp_mean_num = num_mean;
    hyp_lik_num = num_lik;
    flag_cov.assign(num_cov, false);
    exp_cov.assign(num_cov, false);
    type_cov.assign(num_cov, -1);
    fix_param_cov.assign(num_cov, vector<float>());
    flag_mean.assign(num_mean, false);
    exp_mean.assign(num_mean, false);
    type_mean.assign(num_mean, -1);
    fix_param_mean.assign(num_mean, vector<float>());
    flag_lik.assign(num_lik, false);
    exp_lik.assign(num_lik, false);
    type_lik.assign(num_lik, -1);
    fix_param_lik.assign(num_lik, vector<float>());
    cov_varEM.clear();
    cov_varEM_fix.clear();
}

void c_prior::setup_param(const int kernel_index, const vector<int> &kernel_param, const int &mode,
                          const vector<float> &prior_param)
{
    if (kernel_index != 7) {
        std::cout << "Warning: prior mode is only available for LMCSM kernel now; prior will not be effective"
                  << std::endl;
        return;
    }
    if (mode == 0) {
        std::cout << "mode 0: no regularization" << std::endl;
    } else if (mode == 2) {
        std::cout << "mode 2: setup hierarchical gamma prior" << std::endl;
        setup_hier_gamma_prior(kernel_param, prior_param);
    } else {
        std::cout << "undefined setup mode " << mode << "; no changes" << std::endl;
    }
}

void c_prior::setup_hier_gamma_prior(const vector<int> &kernel_param, const vector<float> &prior_param)
{
    const int Q = kernel_param[0], D = kernel_param[1], R = kernel_param[2];
    // variational state [psi (QDR) | delta (QDR) | phi (QR) | tau (QR)], all 1; fixed
    // (alpha, beta, gamma, d) = 0.5 and eta = prior_param[0] (default 50)
    init_cov_varEM(2 * Q * (D * R + R), 1.0);
    init_cov_varEM_fix(5, 0.5);
    set_cov_varEM_fix_one(prior_param.size() > 0 ? prior_param[0] : 50.0, 4);
    const float lap_scale = prior_param.size() > 1 ? prior_param[1] : 0.5f;
    for (int i = 0; i < hyp_cov_num; i++) {
        if (i < Q * D * R) {  // A: normal(0, psi), psi starts at 1
            flag_cov[i] = true;
            exp_cov[i] = false;
            fix_param_cov[i].assign(2, 0.0f);
            fix_param_cov[i][1] = 1.0f;
            type_cov[i] = 1;
        } else if (i < Q * (D * R + 2)) {  // mu, v: no prior, log-stored
            flag_cov[i] = false;
            exp_cov[i] = true;
        } else {  // kappa: laplace(0, beta_lam) on the exp-transformed value
            flag_cov[i] = true;
            exp_cov[i] = true;
            fix_param_cov[i].assign(2, 0.0f);
            fix_param_cov[i][1] = lap_scale;
            type_cov[i] = 2;
        }
    }
}

vector<double> c_prior::prior_lik_normal(const double &x, const vector<float> &param)
{
    vector<double> out(2);
    out[0] = -1.0 * (x - param[0]) * (x - param[0]) / (2.0 * param[1]) - log(2 * PI * param[1]) / 2.0;
    out[1] = -1.0 * (x - param[0]) / param[1];
    return out;
}

vector<double> c_prior::prior_lik_laplace(const double &x, const vector<float> &param)
{
    vector<double> out(2);
    out[0] = (-1.0 * fabs(x - param[0]) / param[1]) - log(2 * param[1]);
    if (x == param[0]) out[1] = 0.0;
    else out[1] = -1.0 * (x > param[0] ? 1.0 : -1.0) / param[1];
    return out;
}

vector<double> c_prior::prior_lik_kde(const double &x, const vector<float> &param)
{
    // param = [bandwidth, sample_0, sample_1, ...]  (prior/c_prior.cpp:165-194)
    const float bw = param[0];
    const int ns = (int)param.size() - 1;
    double lp = 0.0, dlp = 0.0;
    for (int i = 0; i < ns; i++) {
        const double ds = exp(-0.5 * pow((x - param[i + 1]) / bw, 2.0)) / sqrt(2 * PI);
        lp += ds;
        dlp += (x - param[i + 1]) * ds;
    }
    lp = lp / (float(ns) * bw);
    dlp = -1.0 * dlp / (float(ns) * pow(bw, 3.0f));
    dlp = dlp / lp;
    vector<double> out(2);
    out[0] = log(lp);
    out[1] = dlp;
    return out;
}

vector<double> c_prior::one_lik(int type, const double &x, const vector<float> &param)
{
    switch (type) {
        case 1: return prior_lik_normal(x, param);
        case 2: return prior_lik_laplace(x, param);
        case 3: return prior_lik_kde(x, param);
        default: return vector<double>(2, 0.0);
    }
}

vector<double> c_prior::get_one_lik_cov(const double &x, const int &index) const
{
    return one_lik(type_cov[index], x, fix_param_cov[index]);
}
vector<double> c_prior::get_one_lik_lik(const double &x, const int &index) const
{
    return one_lik(type_lik[index], x, fix_param_lik[index]);
}
vector<double> c_prior::get_one_lik_mean(const double &x, const int &index) const
{
    return one_lik(type_mean[index], x, fix_param_mean[index]);
}

void c_prior::init_test_prior(const int kernel_index, const vector<int> &test_kernel_param,
                              const vector<double> &test_mode_param)
{
    if (kernel_index != 7) {
        std::cout << "Warning: testing prior is only set for LMCSM kernel now; prior will not be effective"
                  << std::endl;
        return;
    }
    std::cout << "Info: setup prior to fix zero A elements" << std::endl;
    const int Q = test_kernel_param[0], D = test_kernel_param[1], R = test_kernel_param[2];
    for (int i = D; i < D + Q * D * R; i++)
        if (test_mode_param[i] == 0.0) {
            flag_cov[i - D] = true;
            type_cov[i - D] = 0;
        }
}

bool c_prior::get_one_prior_flag(const int &index) const
{
    if (index < hyp_lik_num) return flag_lik[index];
    if (index < hyp_lik_num + hyp_cov_num) return flag_cov[index - hyp_lik_num];
    return flag_mean[index - hyp_lik_num - hyp_cov_num];
}

int c_prior::get_one_prior_type(const int &index) const
{
    if (index < hyp_lik_num) return type_lik[index];
    if (index < hyp_lik_num + hyp_cov_num) return type_cov[index - hyp_lik_num];
    return type_mean[index - hyp_lik_num - hyp_cov_num];
}

void c_prior::print_status() const
{
    std::cout << "prior table: cov/mean/lik = " << hyp_cov_num << "/" << hyp_mean_num << "/" << hyp_lik_num
              << "; varEM state " << cov_varEM.size() << ", fixed " << cov_varEM_fix.size() << std::endl;
    for (int i = 0; i < hyp_cov_num; i++)
        if (flag_cov[i]) {
            std::cout << "  cov[" << i << "]: type " << type_cov[i];
            for (size_t k = 0; k < fix_param_cov[i].size(); k++) std::cout << " " << fix_param_cov[i][k];
            std::cout << std::endl;
        }
}

// ------------------------------------------------------------------ prior adjustment
namespace {
void adjust_block(const vector<bool> &flag, const vector<int> &type, const vector<bool> &is_exp,
                  const vector<double> &hyp, int offset, bool skip_type_none, bool flag_grad,
                  double &nlml, vector<double> &dnlml, const c_prior &prior, int which)
{
    for (size_t i = 0; i < hyp.size(); i++) {
        if (!flag[i]) continue;
        if (type[i] == 0) {  // clamped: keep the value, kill the gradient
            if (flag_grad) dnlml[i + offset] = 0.0;
            continue;
        }
        if (skip_type_none && type[i] == -1) continue;
        const vector<double> lik = which == 0   ? prior.get_one_lik_lik(hyp[i], (int)i)
                                   : which == 1 ? prior.get_one_lik_cov(hyp[i], (int)i)
                                                : prior.get_one_lik_mean(hyp[i], (int)i);
        nlml -= lik[0];
        if (flag_grad) dnlml[i + offset] -= is_exp[i] ? hyp[i] * lik[1] : lik[1];
    }
}
}  // namespace

void medgp_apply_prior(const c_prior &prior, const vector<double> &lik_hyp, const vector<double> &cov_hyp,
                       const vector<double> &mean_hyp, bool flag_grad, double &nlml,
                       vector<double> &dnlml)
{
    int offset = 0;
    adjust_block(prior.flag_lik, prior.type_lik, prior.exp_lik, lik_hyp, offset, false, flag_grad, nlml, dnlml, prior, 0);
    offset += (int)lik_hyp.size();
    // only the covariance block treats type -1 as "leave untouched" (c_inference_prior.cpp:106-108)
    adjust_block(prior.flag_cov, prior.type_cov, prior.exp_cov, cov_hyp, offset, true, flag_grad, nlml, dnlml, prior, 1);
    offset += (int)cov_hyp.size();
    adjust_block(prior.flag_mean, prior.type_mean, prior.exp_mean, mean_hyp, offset, false, flag_grad, nlml, dnlml, prior, 2);
}
