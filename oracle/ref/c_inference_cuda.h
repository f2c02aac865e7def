// c_inference_cuda.h -- REFERENCE-SIDE BINDING of libmedgp_cuda.so (include/medgp_cuda.h).
//
// This is the class a MedGP maintainer adds to the reference tree: it
// derives from the reference's c_inference (medgpc/src/inference/c_inference.h:38-52), has the
// signature of c_inference_prior::compute_nlml (inference/c_inference_prior.cpp:25-153), does
// the exact-inference part (inference/c_inference_exact.cpp:29-244 incl. the kernel's Gram matrix
// and gradients, kernel/c_kernel_LMC_SM.cpp:152-327) on the GPU, and then applies the prior terms
// exactly as c_inference_prior does, through the reference's own c_prior object.
//
// It lives here, in the test infrastructure, because the repo PROVES the binding by compiling it
// against the unmodified reference sources (oracle/ref/build_ref.sh builds
// oracle/_ref/main_one_train_cuda.o and ref_eval_cuda from /root/reference + this class +
// -lmedgp_cuda) and running the result against the reference's own executables
// (tests/test_gpu_frontends.py, tests/test_host_logic.py).
//
// Compiling a reference translation unit with
//     -include c_inference_cuda.h -DMEDGP_REPLACE_C_INFERENCE_PRIOR
// makes every `c_inference_prior` that the unit constructs (main_one_train.cpp:103-118,
// main_one_test.cpp:119-142) a c_inference_cuda, without touching the reference's sources.
#ifndef C_INFERENCE_CUDA_H
#define C_INFERENCE_CUDA_H

#include <vector>

#include "inference/c_inference.h"
#include "inference/c_inference_prior.h"  // seen here first, so that the macro below cannot rename its class
#include "medgp_cuda.h"

class c_inference_cuda : public c_inference {
  public:
    c_inference_cuda();
    c_inference_cuda(const int &thread_num);
    ~c_inference_cuda();

    // Contract of c_inference::compute_nlml.  nlml and dnlml ([lik | cov | mean] order, prior
    // terms included) are filled; false when the Cholesky still fails after 10 jitter additions.
    // The out-parameters chol_alpha (K^-1 y), chol_factor_inv (row-major lower L^-1) and beta
    // (y^T alpha) are filled on the calls WITHOUT gradient -- the calls the reference follows with
    // GP_Regression::predict (main_one_test.cpp:386-399, gp_regression.cpp:165-196), which reads
    // them -- through medgp_cuda_export_factors; gradient calls (optimiser iterations, whose
    // GP_Regression object is dropped at once, c_objective_one.cpp:61-79) leave the factor on the
    // GPU.  predict() below is the faster route for new code: no n^2 read-back.
    bool compute_nlml(const bool &flag_grad, const vector<int> &meta, const vector<float> &x,
                      const vector<float> &y, c_kernel *kernel, c_meanfunc *meanfunc,
                      c_likelihood *likfunc, c_prior *prior, float *&chol_alpha,
                      float *&chol_factor_inv, float &beta, double &nlml, vector<double> &dnlml);

    // Predictive mean / variance (noise of the test feature included) at (meta2, x2) for the
    // training set and hyper-parameters of the last successful compute_nlml.
    bool predict(const vector<int> &meta2, const vector<float> &x2, vector<double> &mean, vector<double> &var);

    int last_status() const { return status_; }  // 0, or the number of jitter additions

  private:
    void bind(c_kernel *kernel);
    void upload(const vector<int> &meta, const vector<float> &x, const vector<float> &y, int order);
    medgp_ctx *ctx_;
    int Q_, D_, R_, sid_, status_, order_;
    vector<int> meta_;
    vector<float> x_, y_;
    vector<double> theta_;
};

#ifdef MEDGP_REPLACE_C_INFERENCE_PRIOR
#define c_inference_prior c_inference_cuda
#endif

#endif  // C_INFERENCE_CUDA_H
