// medgp_host.h -- host-side mirror of the reference's model / inference / optimiser classes for
// the hot path, re-written from scratch on top of the C ABI in include/medgp_cuda.h.
//
// Class and method names, argument meaning and error behaviour follow the reference so that
// main_one_train / main_one_test read like the originals and a maintainer can swap backends:
//   c_hyperparam            medgpc/src/core/c_hyperparam.{h,cpp}
//   c_kernel, c_kernel_LMC_SM   medgpc/src/kernel/c_kernel.h, c_kernel_LMC_SM.{h,cpp}
//   c_likelihood(_gaussianMO)   medgpc/src/likelihoods/
//   c_meanfunc(_zero)           medgpc/src/mean/
//   c_prior                     medgpc/src/prior/c_prior.{h,cpp}
//   c_inference(_exact,_prior)  medgpc/src/inference/
//   GP_Regression               medgpc/src/core/gp_regression.{h,cpp}
//   c_objective(_one)           medgpc/src/util/c_objective{,_one}.{h,cpp}
//   c_optimizer(_scg,_varEM)    medgpc/src/util/c_optimizer{,_scg,_varEM}.{h,cpp}
// What differs, on purpose:
//   * all numerical work (Gram matrix, Cholesky, solves, gradient, prediction) is done by
//     libmedgp_cuda.so in FP64; there is NO CPU implementation behind these classes;
//   * the float* chol_alpha / chol_factor_inv out-parameters of compute_nlml are replaced by an
//     opaque fit handle (series id + theta) that GP_Regression::predict hands back to the GPU;
//   * the optimisers are built on re-entrant steppers (scg_stepper, varem_stepper) so that
//     thousands of optimiser instances can be driven in lock-step by one batched objective
//     (c_objective_batch); c_optimizer_scg::optimize is the same stepper driven one at a time.
#ifndef MEDGP_HOST_H
#define MEDGP_HOST_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/medgp_cuda.h"

#define PI 3.14159265  // medgpc/src/util/global_settings.h:6 (truncated on purpose)

// ------------------------------------------------------------------------------------------
// flat hyper-parameter vector [lik | cov | mean]          (core/c_hyperparam.cpp:99-121)
class c_hyperparam {
  public:
    c_hyperparam() {}
    c_hyperparam(const std::vector<double> &hyp_all, const int &num_cov, const int &num_mean,
                 const int &num_lik);
    c_hyperparam(const std::vector<double> &hyp_cov, const std::vector<double> &hyp_mean,
                 const std::vector<double> &hyp_lik);
    int get_num_hyp_cov() const { return (int)hyp_cov.size(); }
    int get_num_hyp_mean() const { return (int)hyp_mean.size(); }
    int get_num_hyp_lik() const { return (int)hyp_lik.size(); }
    int get_num_hyp_all() const { return (int)(hyp_cov.size() + hyp_mean.size() + hyp_lik.size()); }
    std::vector<double> get_hyp_cov() const { return hyp_cov; }
    std::vector<double> get_hyp_mean() const { return hyp_mean; }
    std::vector<double> get_hyp_lik() const { return hyp_lik; }
    std::vector<double> get_hyp_all() const;
    void set_hyp_cov(const std::vector<double> &v) { hyp_cov = v; }
    void set_hyp_mean(const std::vector<double> &v) { hyp_mean = v; }
    void set_hyp_lik(const std::vector<double> &v) { hyp_lik = v; }
    void set_hyp_all(const std::vector<double> &hyp, const int &num_cov, const int &num_mean,
                     const int &num_lik);

  private:
    std::vector<double> hyp_cov, hyp_mean, hyp_lik;
};

// ------------------------------------------------------------------------------------------
// kernels                                                  (kernel/c_kernel.h:15-105)
class c_kernel {
  public:
    c_kernel() : kernel_hyp_num(0), kernel_grad_thread(-1), kernel_name("c_kernel") {}
    virtual ~c_kernel() {}
    virtual void set_kernel_hyp(const std::vector<double> &) {}
    virtual void set_kernel_param(const std::vector<int> &) {}
    virtual void reset_coregional_matrix(std::vector<std::vector<double> >) {}
    void print_kernel() const;
    void set_kernel_grad_thread(int n) { kernel_grad_thread = n; }  // kept for CLI parity; unused on GPU
    int get_kernel_grad_thread() const { return kernel_grad_thread; }
    int get_kernel_hyp_num() const { return kernel_hyp_num; }
    std::vector<int> get_kernel_param() const { return kernel_param; }
    std::vector<double> get_kernel_hyp() const { return kernel_hyp; }      // TRANSFORMED (exp applied)
    std::vector<double> get_kernel_hyp_raw() const { return kernel_hyp_raw; }  // as stored in theta

  protected:
    std::vector<double> kernel_hyp, kernel_hyp_raw;
    std::vector<int> kernel_param;
    int kernel_hyp_num, kernel_grad_thread;
    std::string kernel_name;
};

// SM-LMC kernel: kernel_param = (Q, D, R); hyp = [A raw | log mu | log v | log kappa]
//                                                          (kernel/c_kernel_LMC_SM.cpp:51-115)
class c_kernel_LMC_SM : public c_kernel {
  public:
    c_kernel_LMC_SM();
    explicit c_kernel_LMC_SM(const std::vector<int> &input_param);
    c_kernel_LMC_SM(const std::vector<int> &input_param, const std::vector<double> &input_hyp);
    void set_kernel_hyp(const std::vector<double> &input_hyp);
    void set_kernel_param(const std::vector<int> &input_param);
    void compute_coregional_matrix();
    void reset_coregional_matrix(std::vector<std::vector<double> > em_B_array) { coregional_matrix = em_B_array; }
    const std::vector<std::vector<double> > &get_coregional_matrix() const { return coregional_matrix; }
    // scalar base kernel and its derivatives (kernel/c_kernel_LMC_SM.cpp:374-391), FP64
    static double compute_k(const double &rsq, const double &mu, const double &v);
    static double compute_km(const double &rsq, const double &mu, const double &v);
    static double compute_kv(const double &rsq, const double &mu, const double &v);

  private:
    std::vector<std::vector<double> > coregional_matrix;  // Q x (D*D), row-major
};

// ------------------------------------------------------------------------------------------
// likelihoods                                              (likelihoods/c_likelihood.cpp:38-43)
class c_likelihood {
  public:
    c_likelihood() : likfunc_hyp_num(0), likfunc_name("c_likelihood") {}
    explicit c_likelihood(std::vector<int> input_param) : likfunc_param(input_param), likfunc_hyp_num(0), likfunc_name("c_likelihood") {}
    virtual ~c_likelihood() {}
    void print_likfunc() const;
    void set_likfunc_hyp(std::vector<double> input_hyp);  // stores exp(hyp)
    void set_likfunc_param(std::vector<int> input_param) { likfunc_param = input_param; }
    std::vector<int> get_likfunc_param() const { return likfunc_param; }
    std::vector<double> get_likfunc_hyp() const { return likfunc_hyp; }          // TRANSFORMED
    std::vector<double> get_likfunc_hyp_raw() const { return likfunc_hyp_raw; }  // log sigma
    int get_likfunc_hyp_num() const { return likfunc_hyp_num; }

  protected:
    std::vector<double> likfunc_hyp, likfunc_hyp_raw;
    std::vector<int> likfunc_param;
    int likfunc_hyp_num;
    std::string likfunc_name;
};

// one noise level per output; sigma^2_{meta[i]} on the diagonal
//                                                (likelihoods/c_likelihood_gaussianMO.cpp:25-65)
class c_likelihood_gaussianMO : public c_likelihood {
  public:
    c_likelihood_gaussianMO();
    explicit c_likelihood_gaussianMO(std::vector<int> input_param);
};

// ------------------------------------------------------------------------------------------
// mean functions: only the zero mean is on the hot path     (dataio/c_experiment.cpp:389-393)
class c_meanfunc {
  public:
    c_meanfunc() : meanfunc_hyp_num(0), meanfunc_name("c_meanfunc") {}
    virtual ~c_meanfunc() {}
    void print_meanfunc() const;
    void set_meanfunc_hyp(std::vector<double> input_hyp) { meanfunc_hyp = input_hyp; }
    void set_meanfunc_param(std::vector<int> input_param) { meanfunc_param = input_param; }
    std::vector<int> get_meanfunc_param() const { return meanfunc_param; }
    std::vector<double> get_meanfunc_hyp() const { return meanfunc_hyp; }
    int get_meanfunc_hyp_num() const { return meanfunc_hyp_num; }

  protected:
    std::vector<double> meanfunc_hyp;
    std::vector<int> meanfunc_param;
    int meanfunc_hyp_num;
    std::string meanfunc_name;
};

class c_meanfunc_zero : public c_meanfunc {
  public:
    c_meanfunc_zero() { meanfunc_name = "c_meanfunc_zero"; meanfunc_hyp_num = 0; }
};

// ------------------------------------------------------------------------------------------
// per-hyper-parameter prior table + variational-EM state    (prior/c_prior.{h,cpp})
// type: -1 none, 0 clamp (gradient forced to 0), 1 normal(mean, VARIANCE), 2 laplace(loc, scale),
//       3 kde(bandwidth, samples...)
class c_prior {
  public:
    c_prior() : hyp_cov_num(0), hyp_mean_num(0), hyp_lik_num(0) {}
    c_prior(int num_cov, int num_mean, int num_lik);
    void initialize_param(int num_cov, int num_mean, int num_lik);
    void setup_param(const int kernel_index, const std::vector<int> &kernel_param, const int &mode,
                     const std::vector<float> &prior_param);
    void setup_hier_gamma_prior(const std::vector<int> &kernel_param, const std::vector<float> &prior_param);

    std::vector<bool> flag_cov, flag_mean, flag_lik;  // prior active?
    std::vector<bool> exp_cov, exp_mean, exp_lik;     // hyper-parameter stored as log?
    std::vector<std::vector<float> > fix_param_cov, fix_param_mean, fix_param_lik;
    std::vector<int> type_cov, type_mean, type_lik;

    std::vector<double> get_one_lik_cov(const double &x, const int &index) const;
    std::vector<double> get_one_lik_lik(const double &x, const int &index) const;
    std::vector<double> get_one_lik_mean(const double &x, const int &index) const;
    static std::vector<double> prior_lik_normal(const double &x, const std::vector<float> &param);
    static std::vector<double> prior_lik_laplace(const double &x, const std::vector<float> &param);
    static std::vector<double> prior_lik_kde(const double &x, const std::vector<float> &param);

    void init_cov_varEM(int n, double v) { cov_varEM.assign(n, v); }
    void init_cov_varEM_fix(int n, double v) { cov_varEM_fix.assign(n, v); }
    void set_cov_varEM_all(const std::vector<double> &v) { cov_varEM = v; }
    void set_cov_varEM_one(double value, const int &index) { cov_varEM[index] = value; }
    std::vector<double> get_cov_varEM_all() const { return cov_varEM; }
    double get_cov_varEM_one(const int &index) const { return cov_varEM[index]; }
    void set_cov_varEM_fix_all(const std::vector<double> &v) { cov_varEM_fix = v; }
    void set_cov_varEM_fix_one(double value, const int &index) { cov_varEM_fix[index] = value; }
    std::vector<double> get_cov_varEM_fix_all() const { return cov_varEM_fix; }
    double get_cov_varEM_fix_one(const int &index) const { return cov_varEM_fix[index]; }

    // test time: clamp the A entries that the mode kernel has at exactly 0 (c_prior.cpp:118-140)
    void init_test_prior(const int kernel_index, const std::vector<int> &test_kernel_param,
                         const std::vector<double> &test_mode_param);
    bool get_one_prior_flag(const int &index) const;
    int get_one_prior_type(const int &index) const;
    void print_status() const;

  private:
    static std::vector<double> one_lik(int type, const double &x, const std::vector<float> &param);
    int hyp_cov_num, hyp_mean_num, hyp_lik_num;
    std::vector<double> cov_varEM_fix, cov_varEM;
};

// nlml -= log p(theta), g -= (theta *) dlog p for every active prior; clamp -> g = 0
// (inference/c_inference_prior.cpp:59-150).  lik_hyp / cov_hyp are the TRANSFORMED values.
void medgp_apply_prior(const c_prior &prior, const std::vector<double> &lik_hyp,
                       const std::vector<double> &cov_hyp, const std::vector<double> &mean_hyp,
                       bool flag_grad, double &nlml, std::vector<double> &dnlml);

// ------------------------------------------------------------------------------------------
// GPU backend shared by every host object of the process (one context per GPU in use)
class medgp_backend {
  public:
    // context of device `device` for model shape (Q, D, R); created on first use.  Exits with a
    // message when the CUDA library cannot create a context: there is no CPU fallback.
    static medgp_ctx *context(int Q, int D, int R, int device = 0);
    static void shutdown();
    // uploads (meta, x, y) or returns the cached id when the same arrays were uploaded before
    static int series_for(medgp_ctx *ctx, const std::vector<int> &meta, const std::vector<float> &x,
                          const std::vector<float> &y);
    static void drop_series_cache(medgp_ctx *ctx);
};

// what a successful compute_nlml leaves behind for predict (replaces chol_alpha/chol_factor_inv)
struct medgp_fit {
    medgp_ctx *ctx = nullptr;
    int series_id = -1;
    std::vector<double> theta;  // [lik | cov | mean], stored (log/raw) values
    int status = 0;             // jitter count of the factorisation
};

// ------------------------------------------------------------------------------------------
// inference                                                 (inference/c_inference.h:20-57)
class c_inference {
  public:
    c_inference() : inffunc_name("c_inference"), inf_thread_num(1) {}
    explicit c_inference(const int &thread_num) : inffunc_name("c_inference"), inf_thread_num(thread_num) {}
    virtual ~c_inference() {}
    void print_inffunc() const;
    int get_thread_num() const { return inf_thread_num; }
    // Same contract as the reference: true on success, false when the Cholesky still fails
    // after 10 jitter additions; dnlml is cleared and refilled in [lik | cov | mean] order.
    virtual bool compute_nlml(const bool &flag_grad, const std::vector<int> &meta,
                              const std::vector<float> &x, const std::vector<float> &y,
                              c_kernel *kernel, c_meanfunc *meanfunc, c_likelihood *likfunc,
                              c_prior *prior, medgp_fit &fit, double &nlml,
                              std::vector<double> &dnlml) = 0;

  protected:
    std::string inffunc_name;
    int inf_thread_num;  // kept for CLI compatibility (--thread); the GPU path ignores it
};

class c_inference_exact : public c_inference {  // inference/c_inference_exact.cpp:29-244
  public:
    c_inference_exact() { inffunc_name = "c_inference_exact"; }
    explicit c_inference_exact(const int &thread_num) : c_inference(thread_num) { inffunc_name = "c_inference_exact"; }
    bool compute_nlml(const bool &flag_grad, const std::vector<int> &meta, const std::vector<float> &x,
                      const std::vector<float> &y, c_kernel *kernel, c_meanfunc *meanfunc,
                      c_likelihood *likfunc, c_prior *prior, medgp_fit &fit, double &nlml,
                      std::vector<double> &dnlml);
};

class c_inference_prior : public c_inference {  // inference/c_inference_prior.cpp:25-153
  public:
    c_inference_prior() { inffunc_name = "c_inference_prior"; }
    explicit c_inference_prior(const int &thread_num) : c_inference(thread_num) { inffunc_name = "c_inference_prior"; }
    bool compute_nlml(const bool &flag_grad, const std::vector<int> &meta, const std::vector<float> &x,
                      const std::vector<float> &y, c_kernel *kernel, c_meanfunc *meanfunc,
                      c_likelihood *likfunc, c_prior *prior, medgp_fit &fit, double &nlml,
                      std::vector<double> &dnlml);
};

// ------------------------------------------------------------------------------------------
// model facade                                              (core/gp_regression.{h,cpp})
class GP_Regression {
  public:
    GP_Regression();
    GP_Regression(const int &input_dim, c_kernel *input_kernel, c_meanfunc *input_meanfunc,
                  c_likelihood *input_likfunc, c_inference *input_inffunc, c_prior *input_prior);
    void reset(const int &input_dim, c_kernel *input_kernel, c_meanfunc *input_meanfunc,
               c_likelihood *input_likfunc, c_inference *input_inffunc, c_prior *input_prior);
    int get_dim() const { return dim; }
    bool get_flag_trained() const { return flag_trained; }
    double get_neg_log_mlikelihood() const { return nlm_likelihood; }
    std::vector<double> get_dneg_log_mlikelihood() const { return dnlm_likelihood; }
    void train(const bool &flag_grad, const std::vector<int> &meta, const std::vector<float> &x,
               const std::vector<float> &y);
    // {mean[m], var[m]} as the reference returns them (float), computed in FP64 on the GPU
    std::vector<std::vector<float> > predict(const std::vector<int> &meta, const std::vector<int> &meta2,
                                             const std::vector<float> &x, const std::vector<float> &y,
                                             const std::vector<float> &x2);

  private:
    int dim;
    bool flag_trained;
    double nlm_likelihood;
    std::vector<double> dnlm_likelihood;
    medgp_fit fit;
    c_kernel *kernel;
    c_meanfunc *meanfunc;
    c_likelihood *likfunc;
    c_inference *inffunc;
    c_prior *prior;
};

// ------------------------------------------------------------------------------------------
// objectives                                                (util/c_objective{,_one}.{h,cpp})
class c_objective {
  public:
    c_objective() : objective_name("c_objective"), dist_thread_num(1) {}
    virtual ~c_objective() {}
    void print_objective() const;
    virtual void set_dist_thread_num(int) {}
    virtual bool compute_objective(const bool &flag_grad, const std::vector<double> &input_parameter,
                                   double &objective_value, std::vector<double> &gradients,
                                   c_kernel *&input_kernel, c_meanfunc *&input_meanfunc,
                                   c_likelihood *&input_likfunc, c_inference *&input_inffunc,
                                   c_prior *&input_prior) = 0;

  protected:
    std::string objective_name;
    int dist_thread_num;
};

class c_objective_one : public c_objective {
  public:
    c_objective_one();
    c_objective_one(const int &kernel_idx, const std::vector<int> &kernel_param,
                    const std::vector<int> &meta, const std::vector<float> &x,
                    const std::vector<float> &y);
    void set_dist_thread_num(int) { dist_thread_num = 1; }
    bool compute_objective(const bool &flag_grad, const std::vector<double> &input_parameter,
                           double &objective_value, std::vector<double> &gradients,
                           c_kernel *&input_kernel, c_meanfunc *&input_meanfunc,
                           c_likelihood *&input_likfunc, c_inference *&input_inffunc,
                           c_prior *&input_prior);
    const std::vector<int> &meta() const { return obj_meta; }
    const std::vector<float> &x() const { return obj_x; }
    const std::vector<float> &y() const { return obj_y; }

  private:
    std::vector<int> obj_meta;
    std::vector<float> obj_x, obj_y;
    int obj_kernel_idx;
    std::vector<int> obj_kernel_param;
};

// Batched objective: many (series, theta, prior table) evaluated by ONE library call.  This is
// the unit the cohort driver submits every optimiser super-step.
struct medgp_eval_request {
    int series_id;
    const std::vector<double> *theta;
    const c_prior *prior;  // may be null
};
struct medgp_eval_result {
    bool ok;
    double value;
    std::vector<double> grad;
    int status;
};
class c_objective_batch {
  public:
    c_objective_batch(medgp_ctx *ctx, int Q, int D, int R)
        : ctx_(ctx), Q_(Q), D_(D), R_(R), cap_(0), h_theta_(nullptr), h_nlml_(nullptr), h_grad_(nullptr), h_status_(nullptr) {}
    ~c_objective_batch() { release(); }
    c_objective_batch(const c_objective_batch &) = delete;
    c_objective_batch &operator=(const c_objective_batch &) = delete;
    // evaluates every request (NLML + prior terms, gradient when flag_grad); never throws
    void compute(const bool &flag_grad, const std::vector<medgp_eval_request> &reqs,
                 std::vector<medgp_eval_result> &out);
    // frees the page-locked staging buffers; call before the backend context is shut down
    void release();

  private:
    medgp_ctx *ctx_;
    int Q_, D_, R_;
    // theta / results live in page-locked memory that the library copies to and from directly
    size_t cap_;
    double *h_theta_, *h_nlml_, *h_grad_;
    int *h_status_;
};

// ------------------------------------------------------------------------------------------
// optimisers                                  (util/c_optimizer_scg.cpp:25-284, c_optimizer_varEM.cpp)
// Re-entrant form of the reference's "SCG" (Rasmussen's minimize: Polak-Ribiere CG with
// cubic/quadratic line search).  Usage:
//     scg_stepper st(max_iteration, x0);
//     while (st.wants_eval()) { evaluate f,g at st.point(); st.feed(ok, f, g); }
//     st.best_parameter(), st.best_loss()
// Control flow, constants and quirks are those of the reference, evaluation for evaluation.
class scg_stepper {
  public:
    scg_stepper() : state(DONE) {}
    scg_stepper(int max_iteration, const std::vector<double> &init_parameter);
    bool wants_eval() const { return state != DONE; }
    const std::vector<double> &point() const { return probe; }
    void feed(bool ok, double f, const std::vector<double> &g);
    const std::vector<double> &best_parameter() const { return X; }
    double best_loss() const { return fX; }
    int evaluations() const { return n_eval; }

  private:
    enum State { INIT, EXTRAPOLATE, INTERPOLATE, DONE };
    void begin_iteration();
    void begin_extrapolation_pass();
    void request_extrapolation_eval();
    void after_extrapolation_eval();
    void continue_interpolation();
    void finish_iteration();
    void make_probe();
    State state;
    int length, i, n_eval;
    bool ls_failed, obj_flag, success;
    double M;
    double d0, f0_unused, x1, x2, x3, x4, d1, d2, d3, d4, f1, f2, f3, f4, F0;
    double fX;
    std::vector<double> X, X0, s, df0, df3, dF0, probe;
};

class c_optimizer {
  public:
    c_optimizer() : optimizer_name("c_optimizer") {}
    virtual ~c_optimizer() {}
    void print_optimizer() const;
    virtual void optimize(const int &max_iteration, const std::vector<double> &init_parameter,
                          c_objective *objfunc, const bool &display, double &opt_loss,
                          std::vector<double> &opt_parameter, c_kernel *&input_kernel,
                          c_meanfunc *&input_meanfunc, c_likelihood *&input_likfunc,
                          c_inference *&input_inffunc, c_prior *&input_prior) = 0;

  protected:
    std::string optimizer_name;
};

class c_optimizer_scg : public c_optimizer {
  public:
    c_optimizer_scg() { optimizer_name = "c_optimizer_scg"; }
    void optimize(const int &max_iteration, const std::vector<double> &init_parameter,
                  c_objective *objfunc, const bool &display, double &opt_loss,
                  std::vector<double> &opt_parameter, c_kernel *&input_kernel,
                  c_meanfunc *&input_meanfunc, c_likelihood *&input_likfunc,
                  c_inference *&input_inffunc, c_prior *&input_prior);
};

#define DEFAULT_SCG_MAX_ITER 80

// Outer loop of the variational EM for the hierarchical-gamma sparse prior on A
// (util/c_optimizer_varEM.cpp:60-162), without the inner optimiser: budget() is the SCG budget
// of the coming round (100 evaluations for the first 5 rounds, then sub_opt_iter, as a negative
// max_iteration), start() the point it starts from, finish_round() takes the inner optimiser's
// result, applies the 0.5 % early stop and the closed-form tau, phi, delta, psi updates, prunes
// the A entries whose psi hits 0, and reports whether another round follows.  It owns nothing:
// it mutates the c_prior it is given, exactly as the reference does.  Whoever runs the inner
// SCG -- the host stepper below or the device-resident session (medgp_cuda_scg_*) -- shares it.
class varem_rounds {
  public:
    varem_rounds() : prior(nullptr), done_(true) {}
    varem_rounds(int max_iteration, const std::vector<double> &init_parameter, int sub_opt_iter,
                 const std::vector<int> &kernel_param, int lik_num, c_prior *prior);
    bool done() const { return done_; }
    int budget() const { return -(iter < 5 ? 100 : sub_opt_iter); }
    const std::vector<double> &start() const { return opt_parameter; }
    void finish_round(double loss, const std::vector<double> &parameter);
    const std::vector<double> &best_parameter() const { return opt_parameter; }
    double best_loss() const { return opt_loss; }
    int rounds() const { return iter; }

  private:
    c_prior *prior;
    bool done_;
    int max_iter, iter, sub_opt_iter, Q, D, R, lik_num;
    double opt_loss, best_loss_;
    std::vector<double> opt_parameter;
};

// Re-entrant variational EM: varem_rounds around the host SCG stepper.
class varem_stepper {
  public:
    varem_stepper() {}
    varem_stepper(int max_iteration, const std::vector<double> &init_parameter, int sub_opt_iter,
                  const std::vector<int> &kernel_param, int lik_num, c_prior *prior);
    bool wants_eval() const { return !outer.done(); }
    const std::vector<double> &point() const { return scg.point(); }
    void feed(bool ok, double f, const std::vector<double> &g);
    const std::vector<double> &best_parameter() const { return outer.best_parameter(); }
    double best_loss() const { return outer.best_loss(); }
    int rounds() const { return outer.rounds(); }

  private:
    scg_stepper scg;
    varem_rounds outer;
};

// ------------------------------------------------------------------------------------------
// Device-resident lock-step optimisation of many instances (one per patient): the SCG line
// searches run inside libmedgp_cuda.so (medgp_cuda_scg_*: state in HBM, theta never crosses
// PCIe); this driver owns what the reference keeps on the host between SCG runs -- the
// variational-EM rounds and their prior-table updates (c_optimizer_varEM.cpp:60-162).
struct medgp_opt_instance {
    int series_id = -1;
    std::vector<double> init_parameter;
    c_prior *prior = nullptr;     // may be null (no prior terms)
    bool use_varem = false;       // prior mode 2: variational EM around SCG; else plain SCG
    int max_iteration = 0;        // as passed to c_optimizer_*::optimize (negative: evaluation budget / EM rounds)
    int sub_opt_iter = DEFAULT_SCG_MAX_ITER;
    // results
    std::vector<double> opt_parameter;
    double opt_loss = 0.0;
    long evals = 0;
};
// External objective for tests (drives the device state machine through the session's taps
// instead of the GPU evaluation): returns ok, fills f and g for instance `index` at point x.
typedef bool (*medgp_external_objective)(int index, const std::vector<double> &x, double &f, std::vector<double> &g, void *user);
// true when every instance's prior table can run on the device (no kde prior)
bool medgp_device_optimizer_supports(const std::vector<medgp_opt_instance> &inst);
// Runs all instances to completion; returns the number of super-steps.  poll_every: super-steps
// enqueued between two "anyone left?" polls.
long medgp_optimize_on_device(medgp_ctx *ctx, const std::vector<int> &kernel_param, int lik_num,
                              std::vector<medgp_opt_instance> &inst, int poll_every = 8,
                              medgp_external_objective external = nullptr, void *user = nullptr);

class c_optimizer_varEM : public c_optimizer {
  public:
    c_optimizer_varEM() : sub_opt_iter(DEFAULT_SCG_MAX_ITER) { optimizer_name = "c_optimizer_varEM"; }
    void optimize(const int &max_iteration, const std::vector<double> &init_parameter,
                  c_objective *objfunc, const bool &display, double &opt_loss,
                  std::vector<double> &opt_parameter, c_kernel *&input_kernel,
                  c_meanfunc *&input_meanfunc, c_likelihood *&input_likfunc,
                  c_inference *&input_inffunc, c_prior *&input_prior);
    static double update_tau(const float &gamma, const float &d, const float &eta, const double &phi);
    static double update_phi(const int &D, const float &beta, const float &gamma,
                             const double &delta_sum, const double &tau);
    static double update_delta(const float &alpha, const float &beta, const double &psi, const double &phi);
    static double update_psi(const float &alpha, const double &a, const double &delta);
    void set_sub_opt_iter(const int &opt_iter) { sub_opt_iter = opt_iter; }

  private:
    int sub_opt_iter;
};

#endif  // MEDGP_HOST_H
