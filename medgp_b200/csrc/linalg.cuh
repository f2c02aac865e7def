// linalg.cuh -- kernel (2) of the hot path: batched blocked FP64 Cholesky, triangular inverse,
// K^-1 = L^-T L^-1, triangular solves and log-determinant.  Replaces LAPACKE_spotrf /
// spotrs / strtri and the two cblas_sgemm calls that build W in the reference
// (medgpc/src/inference/c_inference_exact.cpp:97-125,130,168-172) -- in FP64 (DESIGN.md
// section 2 explains why FP64 although the reference stores float).
//
// Every heavy product has the single form C(64x64) = sum_l A_l B_l^T over 64x64 column-major
// tiles (common.cuh: gemm_nt_tiles, DMMA.8x8x4 fed by cp.async.bulk):
//   potrf  (left-looking)  L_ik = (K_ik - sum_{l<k} L_il L_kl^T) X_kk^T        X_kk = inv(L_kk)
//   trtri  (row i)         U_ji = -(sum_{l=j}^{i-1} U_jl L_il^T) X_ii^T        U = (L^-1)^T, U_jj = X_jj^T
//   lauum                  (K^-1)_ij = sum_{l>=i} U_il U_jl^T
#pragma once
#include "common.cuh"

__device__ __forceinline__ double *tile_ptr(double *M, int ld, int ti, int tj)
{
    return M + (size_t)tj * MEDGP_NB * ld + (size_t)ti * MEDGP_NB;
}

// second product of the panel kernels: acc = sP * X^T with sP, sX pitch-SLD tiles in smem,
// sP[c][m] (column-major), sX[c][n] = X(n,c)
__device__ __forceinline__ void gemm2_smem(double (&acc)[4][4][2], const double *sP,
                                           const double *sX)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    acc_zero(acc);
    mma_panels(acc, sP, sX, MEDGP_NB / 4, warp & 1, warp >> 1, lane);
}

// ------------------------------------------------------------------ potrf: diagonal block k
// Fused Cholesky + triangular inverse of one 64x64 block by Gauss-Jordan-style elimination on
// the augmented matrix [D | I], entirely in registers.  256 threads: thread (r, g) (r = tid&63,
// g = tid>>6, warp-uniform) owns row r, columns c = 4s+g, s = 0..15, in W[s].
// Step j (pivot d = D_jj after earlier updates, a_r = D_rj):
//      L_rj = a_r / sqrt(d)                       (saved to sL by the owner of column j)
//      W_rc -= (a_r / d) * row_j[c]    for every r > j, c <= r, c != j
//      W_rj  = -a_r / d
// where row_j[c] = D_cj for c > j (symmetry) and the running Y_jc for c < j, Y = rows of
// L^-1 before the final scaling X_rc = Y_rc / L_rr.  One block barrier per step; row_j travels
// through a double-buffered 64-entry shared vector laid out by column group.
#define MEDGP_DIAG_THREADS 256

__device__ __forceinline__ void potf2_inv_gj(double (&W)[16], int r, int g, double *rowbuf /*2x64*/,
                                             double *sL /*pitch SLD*/, int *s_fail)
{
#pragma unroll
    for (int j = 0; j < MEDGP_NB; j++) {
        const int gj = j & 3, sj = j >> 2;
        double *rb = rowbuf + (j & 1) * MEDGP_NB;
        // publish row j: column-j owners write D_rj (r >= j) at the slot of column r ...
        if (g == gj && r >= j) rb[(r & 3) * 16 + (r >> 2)] = W[sj];
        // ... and the four owners of row j write Y_jc, c < j
        if (r == j) {
#pragma unroll
            for (int s = 0; s < 16; s++)
                if (4 * s + g < j) rb[g * 16 + s] = W[s];
        }
        __syncthreads();
        if (r >= j) {
            double d = rb[gj * 16 + sj];
            if (!(d > 0.0)) {  // LAPACK potrf: info > 0 (also catches NaN)
                *s_fail = 1;
                d = 1.0;
            }
            const double ar = rb[(r & 3) * 16 + (r >> 2)];
            if (g == gj) sL[j * MEDGP_SLD + r] = ar * rsqrt(d);  // L_rj (r == j: sqrt(d))
            if (r > j) {
                const double ard = ar / d;
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const int c = 4 * s + g;
                    if (c <= r) {
                        if (s == sj && g == gj) W[s] = -ard;
                        else W[s] = fma(-ard, rb[g * 16 + s], W[s]);
                    }
                }
            }
        }
    }
}

// One CTA per evaluation: D = K_kk - sum_{l<k} L_kl L_kl^T ; L_kk = chol(D) ; X_kk = inv(L_kk).
// Writes L_kk (lower part of the tile), dinv[k], dinvT[k], blk[k] = sum log diag(L_kk) and
// raises fail[] when a pivot is not positive.
__global__ void __launch_bounds__(MEDGP_DIAG_THREADS, 2)
k_potrf_diag(const EvalDesc *__restrict__ descs, int k, int *__restrict__ fail)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    __shared__ double rowbuf[2 * MEDGP_NB];
    __shared__ int s_fail;
    const EvalDesc &e = descs[blockIdx.x];
    if (k >= e.T) return;
    const int ld = e.npad, tid = threadIdx.x;
    gemm_bars_init(&bars);
    if (tid == 0) s_fail = 0;
    double *M = e.M;
    double *sD = smem, *sL = smem + kTileElems;
    if (tid < MEDGP_GEMM_THREADS) {
        double acc[4][4][2];
        acc_zero(acc);
        gemm_nt_tiles(acc, k,
                      [&](int l, const double *&A, int &lda, const double *&B, int &ldb) {
                          A = tile_ptr(M, ld, k, l);
                          B = A;
                          lda = ldb = ld;
                      },
                      smem, &bars);
        // the 4 GEMM warps must all be done with the ring before it is reused as sD
        asm volatile("bar.sync 1, %0;" ::"n"(MEDGP_GEMM_THREADS));
        acc_to_smem(acc, sD, 1.0);
    }
    __syncthreads();
    const int r = tid & 63, g = tid >> 6;
    const double *Kkk = tile_ptr(M, ld, k, k);
    double W[16];
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int c = 4 * s + g;
        W[s] = (c <= r) ? Kkk[(size_t)c * ld + r] - sD[c * MEDGP_SLD + r] : 0.0;
    }
    __syncthreads();  // sD is dead from here on: it becomes the staging tile for X
    potf2_inv_gj(W, r, g, rowbuf, sL, &s_fail);
    __syncthreads();
    const double lrr_inv = 1.0 / sL[r * MEDGP_SLD + r];
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int c = 4 * s + g;
        const double x = (c < r) ? W[s] * lrr_inv : (c == r ? lrr_inv : 0.0);
        sD[c * MEDGP_SLD + r] = x;  // X(r, c)
    }
    __syncthreads();
    // write back: L_kk (lower), dinv (X column-major), dinvT (X^T column-major)
    double *Lkk = tile_ptr(M, ld, k, k);
    double *Xk = e.dinv + (size_t)k * MEDGP_NB * MEDGP_NB;
    double *XTk = e.dinvT + (size_t)k * MEDGP_NB * MEDGP_NB;
    for (int idx = tid; idx < MEDGP_NB * MEDGP_NB; idx += blockDim.x) {
        const int c = idx >> 6, rr = idx & 63;
        if (rr >= c) Lkk[(size_t)c * ld + rr] = sL[c * MEDGP_SLD + rr];
        Xk[c * MEDGP_NB + rr] = sD[c * MEDGP_SLD + rr];   // X(rr, c)
        XTk[c * MEDGP_NB + rr] = sD[rr * MEDGP_SLD + c];  // X^T(rr, c) = X(c, rr)
    }
    if (tid < 32) {
        double s = log(sL[tid * MEDGP_SLD + tid]) + log(sL[(tid + 32) * MEDGP_SLD + tid + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) {
            e.blk[k] = s;
            if (s_fail) fail[e.out_index] = 1;
        }
    }
}

// ------------------------------------------------------------------ potrf: panel below block k
// grid (row tiles i > k, evaluations): L_ik = (K_ik - sum_{l<k} L_il L_kl^T) X_kk^T
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_potrf_panel(const EvalDesc *__restrict__ descs, int k)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    const int i = k + 1 + blockIdx.x;
    if (i >= e.T) return;
    const int ld = e.npad;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    gemm_nt_tiles(acc, k,
                  [&](int l, const double *&A, int &lda, const double *&B, int &ldb) {
                      A = tile_ptr(M, ld, i, l);
                      B = tile_ptr(M, ld, k, l);
                      lda = ldb = ld;
                  },
                  smem, &bars);
    double *Tik = tile_ptr(M, ld, i, k);
    acc_rsub_global(acc, Tik, ld);
    __syncthreads();  // every warp is done with the pipeline buffers
    double *sP = smem, *sX = smem + kTileElems;
    acc_to_smem(acc, sP, 1.0);
    tile_g2s_plain(sX, e.dinv + (size_t)k * MEDGP_NB * MEDGP_NB, MEDGP_NB);
    __syncthreads();
    gemm2_smem(acc, sP, sX);
    acc_to_global(acc, Tik, ld);
}

// ------------------------------------------------------------------ trtri: block row i of L^-1
// grid (j < i, evaluations): U_ji = -(sum_{l=j}^{i-1} U_jl L_il^T) X_ii^T, U_jj = X_jj^T
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_trtri_row(const EvalDesc *__restrict__ descs, int i)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    const int j = blockIdx.x;
    if (i >= e.T || j >= i) return;
    const int ld = e.npad;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    const double *XTj = e.dinvT + (size_t)j * MEDGP_NB * MEDGP_NB;
    gemm_nt_tiles(acc, i - j,
                  [&](int l0, const double *&A, int &lda, const double *&B, int &ldb) {
                      const int l = j + l0;
                      if (l0 == 0) { A = XTj; lda = MEDGP_NB; }
                      else { A = tile_ptr(M, ld, j, l); lda = ld; }
                      B = tile_ptr(M, ld, i, l);
                      ldb = ld;
                  },
                  smem, &bars);
    __syncthreads();
    double *sP = smem, *sX = smem + kTileElems;
    acc_to_smem(acc, sP, -1.0);
    tile_g2s_plain(sX, e.dinv + (size_t)i * MEDGP_NB * MEDGP_NB, MEDGP_NB);
    __syncthreads();
    gemm2_smem(acc, sP, sX);
    acc_to_global(acc, tile_ptr(M, ld, j, i), ld);
}

// lower-triangle tile enumeration: p -> (ti, tj), ti >= tj, row by row
__device__ __forceinline__ void tri_index(int p, int &ti, int &tj)
{
    int i = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= p) i++;
    while (i * (i + 1) / 2 > p) i--;
    ti = i;
    tj = p - i * (i + 1) / 2;
}

// ------------------------------------------------------------------ lauum: K^-1 lower tiles
// grid (lower tiles, evaluations): (K^-1)_ij = sum_{l>=i} U_il U_jl^T  -> written over L_ij
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_lauum(const EvalDesc *__restrict__ descs)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    int i, j;
    tri_index(blockIdx.x, i, j);
    if (i >= e.T) return;
    const int ld = e.npad;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    const double *XTi = e.dinvT + (size_t)i * MEDGP_NB * MEDGP_NB;
    gemm_nt_tiles(acc, e.T - i,
                  [&](int l0, const double *&A, int &lda, const double *&B, int &ldb) {
                      const int l = i + l0;
                      if (l0 == 0) { A = XTi; lda = MEDGP_NB; }
                      else { A = tile_ptr(M, ld, i, l); lda = ld; }
                      if (l == j) { B = XTi; ldb = MEDGP_NB; }  // only when i == j == l
                      else { B = tile_ptr(M, ld, j, l); ldb = ld; }
                  },
                  smem, &bars);
    acc_to_global(acc, tile_ptr(M, ld, i, j), ld);
}

// ------------------------------------------------------------------ forward solves + NLML
// One CTA per evaluation.  rhs rows r (0..nrhs): v_r = L^-1 rhs_r by blocked forward
// substitution (diagonal blocks through X_kk).  Then, for row 0 (= y):
//   nlml = 1/2 z^T z + sum log L_ii + n log(2 PI)/2     (c_inference_exact.cpp:118-120,146-152)
template <int NR>
__device__ __forceinline__ void fwd_solve_group(const EvalDesc &e, double *rhs0, double *sz)
{
    const int ld = e.npad, tid = threadIdx.x, T = e.T;
    for (int k = 0; k < T; k++) {
        // z_k = X_kk * rhs_k : 256 threads = 64 rows x 4 column slices
        const double *Xk = e.dinv + (size_t)k * MEDGP_NB * MEDGP_NB;
        const int r = tid & 63, sl = tid >> 6;
        double part[NR];
#pragma unroll
        for (int q = 0; q < NR; q++) part[q] = 0.0;
        for (int c = sl * 16; c < sl * 16 + 16; c++) {
            if (c > r) break;
            const double xv = Xk[c * MEDGP_NB + r];
#pragma unroll
            for (int q = 0; q < NR; q++) part[q] += xv * rhs0[(size_t)q * ld + k * MEDGP_NB + c];
        }
#pragma unroll
        for (int q = 0; q < NR; q++) sz[(q * 4 + sl) * MEDGP_NB + r] = part[q];
        __syncthreads();
        if (tid < MEDGP_NB) {
#pragma unroll
            for (int q = 0; q < NR; q++) {
                const double z = sz[(q * 4 + 0) * MEDGP_NB + tid] + sz[(q * 4 + 1) * MEDGP_NB + tid] +
                                 sz[(q * 4 + 2) * MEDGP_NB + tid] + sz[(q * 4 + 3) * MEDGP_NB + tid];
                sz[(NR * 4 + q) * MEDGP_NB + tid] = z;
                rhs0[(size_t)q * ld + k * MEDGP_NB + tid] = z;
            }
        }
        __syncthreads();
        // rows below: rhs_i -= sum_c L(i, 64k + c) z_c
        const double *Lcol = e.M + (size_t)k * MEDGP_NB * ld;
        for (int i = (k + 1) * MEDGP_NB + tid; i < ld; i += blockDim.x) {
            double s[NR];
#pragma unroll
            for (int q = 0; q < NR; q++) s[q] = 0.0;
#pragma unroll 8
            for (int c = 0; c < MEDGP_NB; c++) {
                const double lv = Lcol[(size_t)c * ld + i];
#pragma unroll
                for (int q = 0; q < NR; q++) s[q] += lv * sz[(NR * 4 + q) * MEDGP_NB + c];
            }
#pragma unroll
            for (int q = 0; q < NR; q++) rhs0[(size_t)q * ld + i] -= s[q];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_solve(const EvalDesc *__restrict__ descs, ModelDims md, double *__restrict__ out_nlml,
        int *__restrict__ out_status, const int *__restrict__ fail)
{
    __shared__ double sz[(4 * 4 + 4) * MEDGP_NB];
    __shared__ double scratch[64];
    const EvalDesc &e = descs[blockIdx.x];
    const int ld = e.npad, tid = threadIdx.x;
    // rhs row 0 <- y (pad rows 0)
    for (int i = tid; i < ld; i += blockDim.x) e.rhs[i] = (i < e.n) ? e.y[i] : 0.0;
    __syncthreads();
    int r0 = 0;
    while (r0 < e.nrhs) {
        const int g = min(4, e.nrhs - r0);
        double *base = e.rhs + (size_t)r0 * ld;
        if (g == 4) fwd_solve_group<4>(e, base, sz);
        else if (g == 3) fwd_solve_group<3>(e, base, sz);
        else if (g == 2) fwd_solve_group<2>(e, base, sz);
        else fwd_solve_group<1>(e, base, sz);
        r0 += g;
    }
    double v[2] = {0.0, 0.0};
    for (int i = tid; i < ld; i += blockDim.x) v[0] += e.rhs[i] * e.rhs[i];
    for (int k = tid; k < e.T; k += blockDim.x) v[1] += e.blk[k];
    block_reduce_sum<2>(v, scratch);
    if (tid == 0) {
        const bool bad = fail[e.out_index] != 0;
        const double nlml = 0.5 * v[0] + v[1] + e.n * log(2.0 * md.pi) / 2.0;
        out_nlml[e.out_index] = bad ? __longlong_as_double(0x7ff8000000000000LL) : nlml;
        out_status[e.out_index] = bad ? -1 : e.jitter;
    }
}

// ------------------------------------------------------------------ alpha = L^-T z = U z
// grid (row blocks, evaluations), 256 threads = 64 rows x 4 column slices
__global__ void __launch_bounds__(256)
k_alpha(const EvalDesc *__restrict__ descs)
{
    __shared__ double sp[4 * MEDGP_NB];
    const EvalDesc &e = descs[blockIdx.y];
    const int j = blockIdx.x;
    if (j >= e.T) return;
    const int ld = e.npad, tid = threadIdx.x, r = tid & 63, sl = tid >> 6;
    const double *z = e.rhs;
    double s = 0.0;
    // diagonal block: U_jj = X_jj^T
    const double *XT = e.dinvT + (size_t)j * MEDGP_NB * MEDGP_NB;
    for (int c = sl * 16; c < sl * 16 + 16; c++)
        if (c >= r) s += XT[c * MEDGP_NB + r] * z[j * MEDGP_NB + c];
    // strictly upper tiles (j, l), l > j : columns split over the 4 slices
    const double *Urow = e.M + (size_t)j * MEDGP_NB + r;
    for (int c = (j + 1) * MEDGP_NB + sl; c < ld; c += 4) s += Urow[(size_t)c * ld] * z[c];
    sp[sl * MEDGP_NB + r] = s;
    __syncthreads();
    if (tid < MEDGP_NB)
        e.alpha[j * MEDGP_NB + tid] = sp[tid] + sp[MEDGP_NB + tid] + sp[2 * MEDGP_NB + tid] + sp[3 * MEDGP_NB + tid];
}
