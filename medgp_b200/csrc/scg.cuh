// scg.cuh -- device-resident lock-step optimiser: the reference's "SCG" (Rasmussen's minimize,
// Polak-Ribiere conjugate gradients with cubic/quadratic line search,
// medgpc/src/util/c_optimizer_scg.cpp:25-284) as a state machine that lives in HBM, one
// instance per (patient, initialisation), advanced by ONE kernel per optimiser super-step.
//
// A super-step is: evaluate NLML + gradient at every live instance's probe point (the batched
// launch sequence of medgp_cuda.cu, reading theta from and writing nlml/grad/status to the
// session's device arrays), then k_scg_advance -- per instance: add the prior terms
// (inference/c_inference_prior.cpp:59-150), feed (f, g) to the state machine, and write the next
// probe point over the instance's theta.  theta, gradients and line-search state never leave the
// device; the host only polls how many instances still want evaluations.
//
// Control flow, constants and quirks are those of the reference, evaluation for evaluation
// (see medgp_b200/host/optimizer.cpp, the host-side re-entrant form this mirrors, and its
// comments on the reference's quirks): the counter advances by signbit(max_iteration), so only
// negative budgets (function-evaluation counts) are accepted; x1 is always 0 in the cubic
// extrapolation because the reference re-initialises inside its while(1).
// One CTA per instance; every thread carries the scalar state redundantly (all scalars derive
// from block-wide reductions that are broadcast, so all threads take the same branches) and the
// vector work (length P) is strided over the threads.
#pragma once
#include "common.cuh"

#define MEDGP_SCG_THREADS 128

enum { SCG_INIT = 0, SCG_EXTRAPOLATE = 1, SCG_INTERPOLATE = 2, SCG_DONE = 3 };

struct ScgScalars {
    int state, length, i, n_eval;
    int ls_failed, obj_flag, success, pad;
    double M, d0, x1, x2, x3, x4, d1, d2, d3, d4, f1, f2, f3, f4, F0, fX;
};

// per-instance vectors, P doubles each, in this order
enum { SCG_X = 0, SCG_X0, SCG_S, SCG_DF0, SCG_DF3, SCG_DF0BEST /* dF0 */, SCG_NVEC };

struct ScgSession {
    int count, P;
    ScgScalars *sc;        // count
    double *vec;           // count x SCG_NVEC x P
    double *theta;         // count x P: the probe point of every instance (input of the evaluation)
    double *nlml, *grad;   // count, count x P: results of the evaluation
    int *status;           // count: 0 / k jitters / -1 failed
    int *skip;             // count: 1 = instance finished (its evaluation is skipped)
    int *active;           // [0] = instances that still want evaluations after the last advance
    const signed char *ptype;  // count x P prior type per hyper-parameter (-1 none, 0 clamp, 1 normal, 2 laplace) or null
    const signed char *pexp;   // count x P: 1 = the gradient is w.r.t. log(h), so d log p picks up the factor h
    const float *ppar;         // count x P x 2 prior parameters
};

// block-wide sums of NV values, result broadcast to every thread (fixed order: deterministic)
template <int NV>
__device__ __forceinline__ void scg_reduce(double (&v)[NV], double *scratch /* NV * 32 + NV */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) scratch[k * 32 + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = 0.0;
            for (int w = 0; w < nwarp; w++) s += scratch[k * 32 + w];
            scratch[NV * 32 + k] = s;
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = scratch[NV * 32 + k];
    __syncthreads();
}

__device__ __forceinline__ double scg_dot(const double *a, const double *b, int P, double *scratch)
{
    double v[1] = {0.0};
    for (int k = threadIdx.x; k < P; k += blockDim.x) v[0] = fma(a[k], b[k], v[0]);
    scg_reduce<1>(v, scratch);
    return v[0];
}

// Prior terms of one instance (inference/c_inference_prior.cpp:59-150): f -= sum log p(h_i),
// g_i -= (h_i) dlog p(h_i); clamp -> g_i = 0.  h_i is the TRANSFORMED value: exp(theta_i) for the
// noise, mu, v and kappa entries, theta_i itself for A (likelihoods/c_likelihood.cpp:41,
// kernel/c_kernel_LMC_SM.cpp:57-59).  Densities as prior/c_prior.cpp:383-421 (normal with
// VARIANCE p1; laplace with scale p1), float parameters, truncated PI.
__device__ __forceinline__ double scg_apply_prior(const ScgSession &S, int slot, const ModelDims &md, const double *theta,
                                                  double *g, double *scratch)
{
    if (S.ptype == nullptr) return 0.0;
    const signed char *pt = S.ptype + (size_t)slot * S.P, *pe = S.pexp + (size_t)slot * S.P;
    const float *pp = S.ppar + (size_t)slot * S.P * 2;
    const int aLo = md.D, aHi = md.D + md.Q * md.D * md.R;
    double v[1] = {0.0};
    for (int k = threadIdx.x; k < S.P; k += blockDim.x) {
        const int t = pt[k];
        if (t < 0) continue;
        if (t == 0) { g[k] = 0.0; continue; }
        const double h = (k >= aLo && k < aHi) ? theta[k] : exp(theta[k]);
        const double p0 = (double)pp[2 * k], p1 = (double)pp[2 * k + 1];
        double lp, dlp;
        if (t == 1) {
            lp = -1.0 * (h - p0) * (h - p0) / (2.0 * p1) - log(2 * md.pi * p1) / 2.0;
            dlp = -1.0 * (h - p0) / p1;
        } else {
            lp = (-1.0 * fabs(h - p0) / p1) - log(2 * p1);
            dlp = (h == p0) ? 0.0 : -1.0 * (h > p0 ? 1.0 : -1.0) / p1;
        }
        v[0] += lp;
        g[k] -= pe[k] ? h * dlp : dlp;
    }
    scg_reduce<1>(v, scratch);
    return v[0];
}

struct ScgMachine {
    ScgScalars s;
    double *X, *X0, *S, *df0, *df3, *dF0, *probe;
    int P;
    double *scratch;

    static constexpr double INT_ = 0.1, EXT = 3.0, MAXEV = 20, RATIO = 10, SIG = 0.1, RHO = 0.05;

    __device__ __forceinline__ int neg() const { return s.length < 0 ? 1 : 0; }  // signbit(max_iteration)
    __device__ __forceinline__ void copy(double *dst, const double *src) const
    {
        for (int k = threadIdx.x; k < P; k += blockDim.x) dst[k] = src[k];
    }
    __device__ __forceinline__ void make_probe() const
    {
        for (int k = threadIdx.x; k < P; k += blockDim.x) probe[k] = X[k] + s.x3 * S[k];
    }
    __device__ __forceinline__ void steepest()
    {
        for (int k = threadIdx.x; k < P; k += blockDim.x) S[k] = -1.0 * df0[k];
        s.d0 = -1.0 * scg_dot(S, S, P, scratch);
    }

    // phases of the reference's nested loops; each returns the next phase, PH_EVAL = a probe was
    // written and the instance waits for its evaluation
    enum { PH_BEGIN_ITER, PH_BEGIN_EXTRA, PH_REQ_EXTRA, PH_AFTER_EXTRA, PH_CONT_INTERP, PH_FINISH, PH_EVAL };

    __device__ void run(int ph)
    {
        while (ph != PH_EVAL) {
            switch (ph) {
            case PH_BEGIN_ITER:
                if (!(s.i < abs(s.length))) { s.state = SCG_DONE; return; }
                s.i += neg();  // c_optimizer_scg.cpp:88
                copy(X0, X);
                s.F0 = s.fX;
                copy(dF0, df0);
                s.M = (s.length > 0) ? MAXEV : (double)min((int)MAXEV, abs(s.length) - s.i);
                ph = PH_BEGIN_EXTRA;
                break;
            case PH_BEGIN_EXTRA:  // c_optimizer_scg.cpp:101-110: re-initialised on every pass
                s.x2 = 0.0; s.f2 = s.fX; s.d2 = s.d0;
                s.f3 = s.fX;
                copy(df3, df0);
                s.success = 0;
                ph = PH_REQ_EXTRA;
                break;
            case PH_REQ_EXTRA:
                if (!s.success && s.M > 0) {
                    s.M -= 1;
                    s.i += neg();
                    make_probe();
                    s.state = SCG_EXTRAPOLATE;
                    ph = PH_EVAL;
                } else {
                    ph = PH_AFTER_EXTRA;
                }
                break;
            case PH_AFTER_EXTRA: {
                if (s.f3 < s.F0) {
                    for (int k = threadIdx.x; k < P; k += blockDim.x) X0[k] = X[k] + s.x3 * S[k];
                    s.F0 = s.f3;
                    copy(dF0, df3);
                }
                s.d3 = scg_dot(df3, S, P, scratch);
                if ((s.d3 > SIG * s.d0) || (s.f3 > (s.fX + s.x3 * RHO * s.d0)) || (s.M == 0)) { ph = PH_CONT_INTERP; break; }
                s.x1 = s.x2; s.f1 = s.f2; s.d1 = s.d2;
                s.x2 = s.x3; s.f2 = s.f3; s.d2 = s.d3;
                const double A = 6.0 * (s.f1 - s.f2) + 3.0 * (s.d2 + s.d1) * (s.x2 - s.x1);
                const double B = 3.0 * (s.f2 - s.f1) - (2.0 * s.d1 + s.d2) * (s.x2 - s.x1);
                const double temp = B * B - A * s.d1 * (s.x2 - s.x1);
                if (temp < 0) {
                    s.x3 = s.x2 * EXT;
                } else {
                    s.x3 = s.x1 - (s.d1 * ((s.x2 - s.x1) * (s.x2 - s.x1)) / (B + sqrt(temp)));
                    if (isnan(s.x3) || isinf(s.x3) || (s.x3 < 0)) s.x3 = s.x2 * EXT;
                    else if (s.x3 > s.x2 * EXT) s.x3 = s.x2 * EXT;
                    else if (s.x3 < (s.x2 + INT_ * (s.x2 - s.x1))) s.x3 = s.x2 + INT_ * (s.x2 - s.x1);
                }
                ph = PH_BEGIN_EXTRA;
                break;
            }
            case PH_CONT_INTERP:
                if (((fabs(s.d3) > -1.0 * SIG * s.d0) || (s.f3 > (s.fX + s.x3 * RHO * s.d0))) && (s.M > 0)) {
                    if ((s.d3 > 0) || (s.f3 > (s.fX + s.x3 * RHO * s.d0))) { s.x4 = s.x3; s.f4 = s.f3; s.d4 = s.d3; }
                    else { s.x2 = s.x3; s.f2 = s.f3; s.d2 = s.d3; }
                    const double w = s.x4 - s.x2;
                    if (s.f4 > s.fX) {
                        s.x3 = s.x2 - (0.5 * s.d2 * (w * w)) / (s.f4 - s.f2 - s.d2 * w);
                        if (isnan(s.x3) || isinf(s.x3)) s.x3 = (s.x2 + s.x4) / 2.0;
                    } else {
                        const double A = 6.0 * (s.f2 - s.f4) / w + 3.0 * (s.d4 + s.d2);
                        const double B = 3.0 * (s.f4 - s.f2) - (2.0 * s.d2 + s.d4) * w;
                        const double disc = B * B - A * s.d2 * (w * w);
                        if (disc < 0) {
                            s.x3 = (s.x2 + s.x4) / 2.0;
                        } else {
                            s.x3 = s.x2 + (sqrt(disc) - B) / A;
                            if (isnan(s.x3) || isinf(s.x3)) s.x3 = (s.x2 + s.x4) / 2.0;
                        }
                    }
                    s.x3 = fmax(fmin(s.x3, s.x4 - INT_ * (s.x4 - s.x2)), s.x2 + INT_ * (s.x4 - s.x2));
                    make_probe();
                    s.state = SCG_INTERPOLATE;
                    ph = PH_EVAL;
                } else {
                    ph = PH_FINISH;
                }
                break;
            case PH_FINISH:
                if (s.obj_flag && (fabs(s.d3) < -1.0 * SIG * s.d0) && (s.f3 < (s.fX + s.x3 * RHO * s.d0))) {
                    for (int k = threadIdx.x; k < P; k += blockDim.x) X[k] = X[k] + s.x3 * S[k];
                    s.fX = s.f3;
                    double v[3] = {0.0, 0.0, 0.0};  // Polak-Ribiere direction
                    for (int k = threadIdx.x; k < P; k += blockDim.x) {
                        v[0] = fma(df3[k], df3[k], v[0]);
                        v[1] = fma(df3[k], df0[k], v[1]);
                        v[2] = fma(df0[k], df0[k], v[2]);
                    }
                    scg_reduce<3>(v, scratch);
                    const double beta = (v[0] - v[1]) / v[2];
                    for (int k = threadIdx.x; k < P; k += blockDim.x) S[k] = beta * S[k] - df3[k];
                    copy(df0, df3);
                    s.d3 = s.d0;
                    s.d0 = scg_dot(df0, S, P, scratch);
                    if (s.d0 > 0) steepest();
                    s.x3 = s.x3 * fmin(RATIO, s.d3 / (s.d0 - 2.220446049250313e-16));  // pow(2, -52)
                    s.ls_failed = 0;
                } else {
                    copy(X, X0);
                    s.fX = s.F0;
                    copy(df0, dF0);
                    steepest();
                    s.x3 = 1.0 / (1.0 - s.d0);
                    s.ls_failed = 1;
                }
                ph = PH_BEGIN_ITER;
                break;
            }
        }
    }

    // one evaluation result for the instance's current probe (g already carries the prior terms)
    __device__ void feed(bool ok, double f, const double *g)
    {
        s.n_eval++;
        if (s.state == SCG_INIT) {
            if (!ok) {  // the reference would continue with indeterminate values; stop instead
                s.fX = __longlong_as_double(0x7ff8000000000000LL);
                s.state = SCG_DONE;
                return;
            }
            s.fX = f;
            copy(df0, g);
            s.i += neg();
            steepest();
            s.x3 = 1.0 / (1.0 - s.d0);  // red = 1
            run(PH_BEGIN_ITER);
        } else if (s.state == SCG_EXTRAPOLATE) {
            s.obj_flag = ok;
            if (ok) { s.f3 = f; copy(df3, g); }
            if (!ok || isinf(s.f3) || isnan(s.f3)) s.x3 = (s.x2 + s.x3) / 2.0;
            else s.success = 1;
            run(PH_REQ_EXTRA);
        } else if (s.state == SCG_INTERPOLATE) {
            s.obj_flag = ok;
            if (ok) { s.f3 = f; copy(df3, g); }
            if (s.obj_flag && s.f3 < s.F0) {
                copy(X0, probe);
                s.F0 = s.f3;
                copy(dF0, df3);
            }
            s.M -= 1;
            s.i += neg();
            s.d3 = scg_dot(df3, S, P, scratch);
            run(PH_CONT_INTERP);
        }
    }
};

// (re)start every instance at theta0 with budget max_iteration[slot] (negative: evaluations).
// A non-negative budget leaves the instance finished at theta0 (loss NaN) -- see the header.
__global__ void __launch_bounds__(MEDGP_SCG_THREADS)
k_scg_start(ScgSession S, const double *__restrict__ theta0, const int *__restrict__ max_iteration)
{
    const int slot = blockIdx.x;
    double *v = S.vec + (size_t)slot * SCG_NVEC * S.P;
    for (int k = threadIdx.x; k < S.P; k += blockDim.x) {
        const double t = theta0[(size_t)slot * S.P + k];
        v[SCG_X * S.P + k] = t;
        S.theta[(size_t)slot * S.P + k] = t;
    }
    if (threadIdx.x == 0) {
        ScgScalars z;
        memset(&z, 0, sizeof(z));
        z.length = max_iteration[slot];
        z.state = z.length < 0 ? SCG_INIT : SCG_DONE;
        if (z.length >= 0) z.fX = __longlong_as_double(0x7ff8000000000000LL);
        S.sc[slot] = z;
        S.skip[slot] = z.length < 0 ? 0 : 1;
        if (slot == 0) S.active[0] = 0;
    }
}

// counts the instances that want an evaluation (after k_scg_start or k_scg_advance)
__global__ void k_scg_count(ScgSession S)
{
    int n = 0;
    for (int b = threadIdx.x; b < S.count; b += blockDim.x) n += S.skip[b] ? 0 : 1;
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    __shared__ int part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
        S.active[0] = t;
    }
}

// one optimiser step of every live instance: prior terms, feed, next probe
__global__ void __launch_bounds__(MEDGP_SCG_THREADS)
k_scg_advance(ScgSession S, ModelDims md)
{
    __shared__ double scratch[3 * 32 + 3];
    const int slot = blockIdx.x;
    if (S.skip[slot]) return;
    ScgMachine m;
    m.s = S.sc[slot];
    m.P = S.P;
    m.scratch = scratch;
    double *v = S.vec + (size_t)slot * SCG_NVEC * S.P;
    m.X = v + SCG_X * S.P; m.X0 = v + SCG_X0 * S.P; m.S = v + SCG_S * S.P;
    m.df0 = v + SCG_DF0 * S.P; m.df3 = v + SCG_DF3 * S.P; m.dF0 = v + SCG_DF0BEST * S.P;
    m.probe = S.theta + (size_t)slot * S.P;
    double *g = S.grad + (size_t)slot * S.P;
    const bool ok = S.status[slot] >= 0;
    double f = S.nlml[slot];
    if (ok) f -= scg_apply_prior(S, slot, md, m.probe, g, scratch);
    __syncthreads();
    m.feed(ok, f, g);
    __syncthreads();
    if (threadIdx.x == 0) {
        S.sc[slot] = m.s;
        if (m.s.state == SCG_DONE) S.skip[slot] = 1;
    }
}

// before a chunk's launch sequence: retire the descriptors of finished instances
__global__ void k_apply_skip(EvalDesc *__restrict__ descs, int count, const int *__restrict__ ext_skip)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < count && ext_skip[descs[b].out_index]) descs[b].skip = 1;
}
