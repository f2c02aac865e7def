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

__device__ __forceinline__ double *tile_ptr(double *M, int T, int ti, int tj)
{
    return M + tile_off(T, ti, tj);
}

// second product of the panel kernels: acc = sP * X^T with sP, sX pitch-SLD tiles in smem,
// sP[c][m] (column-major), sX[c][n] = X(n,c)
__device__ __forceinline__ void gemm2_smem(double (&acc)[4][4][2], const double *sP,
                                           const double *sX)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wn = warp >> 1;
    acc_zero(acc);
    // X(n, c) = 0 for c > n: output columns 0..31 (wn == 0) only need the first 32 k-steps
    mma_panels(acc, sP, sX, wn == 0 ? MEDGP_NB / 8 : MEDGP_NB / 4, warp & 1, wn, lane);
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

// ------------------------------------------------------------------ fused forward solve
// The triangular solve z = L^-1 rhs rides on the factorisation: the diagonal kernel of step k
// turns the already-updated block rhs_k into z_k = X_kk rhs_k, and every panel CTA (i, k)
// pushes its fresh tile into the rows below, rhs_i -= L_ik z_k.  All right-hand sides of the
// evaluation (row 0 = y, rows 1.. = cross-covariance columns) are carried along.
// sTile: pitch-SLD tile in smem holding the operator (element (r, c) at c*SLD + r);
// 128 threads: r = tid & 63, half = tid >> 6 splits the columns by parity.
__device__ __forceinline__ void tile_matvec_rhs(const EvalDesc &e, const double *sTile, const double *vec_base,
                                                double *out_base, bool lower_only, bool subtract,
                                                double *red /*2*64*/)
{
    const int tid = threadIdx.x, r = tid & 63, half = tid >> 6, ld = e.npad;
    for (int q = 0; q < e.nrhs; q++) {
        const double *v = vec_base + (size_t)q * ld;
        double s = 0.0;
        const int cmax = lower_only ? r : MEDGP_NB - 1;
#pragma unroll 4
        for (int c = half; c <= cmax; c += 2) s += sTile[c * MEDGP_SLD + r] * v[c];
        red[half * MEDGP_NB + r] = s;
        __syncthreads();
        if (half == 0) {
            const double tot = red[r] + red[MEDGP_NB + r];
            double *o = out_base + (size_t)q * ld;
            o[r] = subtract ? o[r] - tot : tot;
        }
        __syncthreads();
    }
}

// rhs_i -= L_ik z_k straight from the accumulator fragments (acc = L_ik): every thread forms the
// partial dot products of its 4 row sub-tiles over its 8 columns, the 4 lanes of a quad are
// combined with shuffles, the two column halves (wn) through a 2x64 shared scratch.
__device__ __forceinline__ void acc_matvec_rhs(const double (&acc)[4][4][2], const EvalDesc &e,
                                               const double *zk_base, double *out_base, double *red /*2*64*/)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, ld = e.npad;
    const int wm = warp & 1, wn = warp >> 1, r = lane >> 2, kq = lane & 3;
    for (int q = 0; q < e.nrhs; q++) {
        const double *z = zk_base + (size_t)q * ld;
        double ps[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int col = wn * 32 + 8 * b + 2 * kq;
            const double z0 = __ldcg(z + col), z1 = __ldcg(z + col + 1);
#pragma unroll
            for (int a = 0; a < 4; a++) ps[a] += acc[a][b][0] * z0 + acc[a][b][1] * z1;
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            ps[a] += __shfl_xor_sync(0xffffffffu, ps[a], 1);
            ps[a] += __shfl_xor_sync(0xffffffffu, ps[a], 2);
        }
        if (kq == 0)
#pragma unroll
            for (int a = 0; a < 4; a++) red[wn * MEDGP_NB + wm * 32 + 8 * a + r] = ps[a];
        __syncthreads();
        if (threadIdx.x < MEDGP_NB) {
            double *o = out_base + (size_t)q * ld;
            o[threadIdx.x] -= red[threadIdx.x] + red[MEDGP_NB + threadIdx.x];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ potrf: diagonal block k
// Fused Cholesky + triangular inverse of one 64x64 block by Gauss-Jordan-style elimination on
// the augmented matrix [D | I], entirely in registers.  128 threads: thread (r, g) (r = tid&63,
// g = tid>>6, warp-uniform) owns row r and the 32 columns of parity g.
// Step j (pivot d = D_jj after earlier updates, a_r = D_rj):
//      L_rj = a_r / sqrt(d)                       (saved to sL by the owner of column j)
//      W_rc -= (a_r / d) * row_j[c]    for every r > j, c != j   (c > r: dead slots, harmless)
//      W_rj  = -a_r / d
// where row_j[c] = D_cj for c > j (symmetry) and the running Y_jc for c < j, Y = rows of
// L^-1 before the final scaling X_rc = Y_rc / L_rr.  One block barrier per step; row_j travels
// through a double-buffered 64-entry shared vector laid out by column parity (16-byte
// broadcast reads).  To keep the code inside the instruction cache only 8 steps are unrolled:
// after each panel of 8 columns the register file is rotated by 4 slots, so the pivot columns
// always sit in W[0..3] and every register index stays static:
//      during panel p, W[pos] holds column c = 2*((pos + 4p) mod 32) + g.
#define MEDGP_DIAG_THREADS 128

// shared scratch: d[2][2*32] (column j of D by parity), y[2][2][64] (row j of Y in the rotated
// slot coordinates of its owner, per parity), rs[2] = 1/sqrt(pivot)
struct GjBufs {
    double d[2][MEDGP_NB];
    double y[2][2][MEDGP_NB];
    double rs[2];
};

__device__ __forceinline__ void potf2_inv_gj(double (&W)[32], int r, int g, GjBufs *gb,
                                             double *sL /*pitch SLD*/, int *s_fail)
{
    const int rslot = (r & 1) * 32 + (r >> 1);
    // 1/sqrt of the NEXT pivot is computed during the previous update sweep (by every thread on
    // its own slot, branch-free, so the compiler interleaves it with the sweep's FMAs; only the
    // owner's value is published).  The first one here.
    double rs_next = rsqrt(W[0]);
    if (r == 0 && g == 0 && !(W[0] > 0.0)) *s_fail = 1;
#pragma unroll 1
    for (int p = 0; p < 8; p++) {
        const int base = 4 * p;  // W[pos] holds column 2*((pos + base) mod 32) + g
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const int j = 8 * p + jj;
            const int gj = jj & 1, pj = jj >> 1;  // static: parity and register slot of column j
            const int pn = (jj + 1) >> 1;         // ... and slot of column j + 1 (4 when jj == 7)
            double *db = gb->d[jj & 1];
            double *yb = gb->y[jj & 1][g];
            if (g == gj && r >= j) db[rslot] = W[pj];  // column j of D: D_rj, r >= j
            if (r == j) {                              // row j of Y (all 32 slots, rotated coords)
                if (g == gj) {
                    gb->rs[jj & 1] = rs_next;
                    if (!(W[pj] > 0.0)) *s_fail = 1;   // LAPACK potrf: info > 0 (also NaN)
                }
#pragma unroll
                for (int pos = 0; pos < 32; pos += 2)
                    *reinterpret_cast<double2 *>(yb + base + pos) = make_double2(W[pos], W[pos + 1]);
            }
            __syncthreads();
            if (r >= j) {
                const double rs = gb->rs[jj & 1];
                const double ar = db[rslot];
                const double l = ar * rs;  // L_rj = a_r / sqrt(d)   (r == j: sqrt(d))
                if (g == gj) sL[j * MEDGP_SLD + r] = l;
                if (r > j) {
                    const double nard = -ar * (rs * rs);  // -a_r / d
                    const double *dsrc = db + g * 32 + base;  // future columns: D_cj
                    const double *ysrc = yb + base;           // past columns:   Y_jc
                    // slots 0..3: the current panel (columns 8p + 2 pos + g), per-slot choice
                    {
                        const double2 d0 = *reinterpret_cast<const double2 *>(dsrc);
                        const double2 d1 = *reinterpret_cast<const double2 *>(dsrc + 2);
                        const double2 y0 = *reinterpret_cast<const double2 *>(ysrc);
                        const double2 y1 = *reinterpret_cast<const double2 *>(ysrc + 2);
                        const double dv[4] = {d0.x, d0.y, d1.x, d1.y};
                        const double yv[4] = {y0.x, y0.y, y1.x, y1.y};
#pragma unroll
                        for (int pos = 0; pos < 4; pos++) {
                            const double v = (2 * pos + g < jj) ? yv[pos] : dv[pos];
                            W[pos] = fma(nard, v, W[pos]);
                        }
                        if (g == gj) W[pj] = nard;
                    }
                    // slot group 1 next: for jj == 7 it holds the next pivot
                    {
                        const double *sp = (1 >= 8 - p) ? ysrc : dsrc;
                        const double2 v0 = *reinterpret_cast<const double2 *>(sp + 4);
                        const double2 v1 = *reinterpret_cast<const double2 *>(sp + 6);
                        W[4] = fma(nard, v0.x, W[4]);
                        W[5] = fma(nard, v0.y, W[5]);
                        W[6] = fma(nard, v1.x, W[6]);
                        W[7] = fma(nard, v1.y, W[7]);
                    }
                    rs_next = rsqrt(W[pn]);  // meaningful in the thread that owns pivot j + 1
#pragma unroll
                    for (int q = 2; q < 8; q++) {
                        const double *sp = (q >= 8 - p) ? ysrc : dsrc;
                        const double2 v0 = *reinterpret_cast<const double2 *>(sp + 4 * q);
                        const double2 v1 = *reinterpret_cast<const double2 *>(sp + 4 * q + 2);
                        W[4 * q] = fma(nard, v0.x, W[4 * q]);
                        W[4 * q + 1] = fma(nard, v0.y, W[4 * q + 1]);
                        W[4 * q + 2] = fma(nard, v1.x, W[4 * q + 2]);
                        W[4 * q + 3] = fma(nard, v1.y, W[4 * q + 3]);
                    }
                }
            }
        }
        // rotate the register file by 4 slots
        const double t0 = W[0], t1 = W[1], t2 = W[2], t3 = W[3];
#pragma unroll
        for (int pos = 0; pos < 28; pos++) W[pos] = W[pos + 4];
        W[28] = t0; W[29] = t1; W[30] = t2; W[31] = t3;
    }
}

// Factor one diagonal block: D = K_kk - C with the accumulated product C in sD (pitch SLD);
// L_kk = chol(D), X_kk = inv(L_kk).  Writes L_kk (lower part of the tile), dinv[k], dinvT[k],
// blk[k] = sum log diag(L_kk), turns rhs block k into z_k = X_kk rhs_k and raises fail[] when a
// pivot is not positive.  128 threads; sD / sL are the two halves of the dynamic smem ring.
__device__ __forceinline__ void diag_block_factor(const EvalDesc &e, int k, double *sD, double *sL,
                                                  GjBufs *gjb, int *s_fail, int *__restrict__ fail,
                                                  bool have_product = true)
{
    const int T = e.T, tid = threadIdx.x;
    const int r = tid & 63, g = tid >> 6;
    const double *Kkk = tile_ptr(e.M, T, k, k);
    double W[32];
#pragma unroll
    for (int s = 0; s < 32; s++) {
        const int c = 2 * s + g;
        W[s] = (c <= r) ? Kkk[c * MEDGP_SLD + r] - (have_product ? sD[c * MEDGP_SLD + r] : 0.0) : 0.0;
    }
    __syncthreads();  // sD is dead from here on: it becomes the staging tile for X
    potf2_inv_gj(W, r, g, gjb, sL, s_fail);
    __syncthreads();
    const double lrr_inv = 1.0 / sL[r * MEDGP_SLD + r];
#pragma unroll
    for (int s = 0; s < 32; s++) {
        const int c = 2 * s + g;
        const double x = (c < r) ? W[s] * lrr_inv : (c == r ? lrr_inv : 0.0);
        sD[c * MEDGP_SLD + r] = x;  // X(r, c)
    }
    __syncthreads();
    // forward solve, block k: z_k = X_kk rhs_k (in place)
    tile_matvec_rhs(e, sD, e.rhs + k * MEDGP_NB, e.rhs + k * MEDGP_NB, true, false, gjb->d[0]);
    // write back: L_kk (lower), dinv (X column-major), dinvT (X^T column-major)
    double *Lkk = tile_ptr(e.M, T, k, k);
    double *Xk = e.dinv + (size_t)k * kTileElems;
    double *XTk = e.dinvT + (size_t)k * kTileElems;
    for (int idx = tid; idx < MEDGP_NB * MEDGP_NB; idx += blockDim.x) {
        const int c = idx >> 6, rr = idx & 63;
        if (rr >= c) Lkk[c * MEDGP_SLD + rr] = sL[c * MEDGP_SLD + rr];
        Xk[c * MEDGP_SLD + rr] = sD[c * MEDGP_SLD + rr];   // X(rr, c)
        XTk[c * MEDGP_SLD + rr] = sD[rr * MEDGP_SLD + c];  // X^T(rr, c) = X(c, rr)
    }
    if (tid < 32) {
        double s = log(sL[tid * MEDGP_SLD + tid]) + log(sL[(tid + 32) * MEDGP_SLD + tid + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) {
            e.blk[k] = s;
            if (*s_fail) fail[e.out_index] = 1;
        }
    }
}

// Stand-alone diagonal kernel, one CTA per evaluation: D = K_kk - sum_{l<depth} L_kl L_kl^T.
// Used for block 0, for the right-looking path (depth 0: the tile is already updated), and as
// the general fallback; in the left-looking path blocks k >= 1 are factored inside the panel
// kernel of step k-1 (see k_potrf_panel).
__global__ void __launch_bounds__(MEDGP_DIAG_THREADS, 3)
k_potrf_diag(const EvalDesc *__restrict__ descs, int k, int depth, int *__restrict__ fail)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    __shared__ __align__(16) GjBufs gjb;
    __shared__ int s_fail;
    const EvalDesc &e = descs[blockIdx.x];
    if (k >= e.T) return;
    const int T = e.T, tid = threadIdx.x;
    gemm_bars_init(&bars);
    if (tid == 0) s_fail = 0;
    double *M = e.M;
    double *sD = smem, *sL = smem + kTileElems;
    {
        double acc[4][4][2];
        acc_zero(acc);
        gemm_nt_tiles(acc, depth,
                      [&](int l, const double *&A, const double *&B) {
                          A = tile_ptr(M, T, k, l);
                          B = A;
                      },
                      smem, &bars);
        __syncthreads();  // all warps are done with the ring before it is reused as sD
        acc_to_smem(acc, sD, 1.0);
    }
    __syncthreads();
    diag_block_factor(e, k, sD, sL, &gjb, &s_fail, fail);
}

// ------------------------------------------------------------------ potrf: panel below block k
// grid (row tiles i > k, evaluations): L_ik = (K_ik - sum_{l<k} L_il L_kl^T) X_kk^T
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_potrf_panel(const EvalDesc *__restrict__ descs, int k, int depth, int fold_diag, int *__restrict__ fail)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    const int i = k + 1 + blockIdx.x;
    if (i >= e.T) return;
    const int T = e.T;
    double *M = e.M;
    double *Tik = tile_ptr(M, T, i, k);
    const double *Xk = e.dinv + (size_t)k * kTileElems;
    __shared__ double red[2 * MEDGP_NB];
    prefetch_tile_l2(Tik);  // epilogue operands: start them towards L2 now
    prefetch_tile_l2(Xk);
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    gemm_nt_tiles(acc, depth,
                  [&](int l, const double *&A, const double *&B) {
                      A = tile_ptr(M, T, i, l);
                      B = tile_ptr(M, T, k, l);
                  },
                  smem, &bars);
    __syncthreads();  // every warp is done with the pipeline buffers
    double *sP = smem, *sX = smem + kTileElems;
    tile_bulk_g2s(sX, Xk, &bars);  // X_kk arrives while P = K_ik - C is formed
    acc_rsub_global(acc, Tik);
    acc_to_smem(acc, sP, 1.0);
    tile_bulk_wait(&bars);
    __syncthreads();
    gemm2_smem(acc, sP, sX);
    acc_to_global(acc, Tik);
    // forward solve: push the fresh tile into the right-hand sides of block row i
    acc_matvec_rhs(acc, e, e.rhs + k * MEDGP_NB, e.rhs + i * MEDGP_NB, red);
    if (fold_diag) {
        // apply this tile to its diagonal block right away, K_ii -= L_ik L_ik^T, so that no
        // diagonal kernel has a k-tile product to do.  With fold_diag == 2 the CTA of row k+1,
        // whose diagonal block is now complete, factors it on the spot: the next step's
        // diagonal kernel disappears and its latency hides behind the other panel CTAs.
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        acc_to_smem(acc, sP, 1.0);  // (acc_matvec_rhs ended with a block barrier: GEMM2 is done with sP)
        __syncthreads();
        acc_zero(acc);
        mma_panels(acc, sP, sP, MEDGP_NB / 4, warp & 1, warp >> 1, lane);
        if (fold_diag == 2 && i == k + 1) {
            __shared__ __align__(16) GjBufs gjb;
            __shared__ int s_fail;
            if (threadIdx.x == 0) s_fail = 0;
            __syncthreads();  // everyone is done reading sP
            acc_to_smem(acc, sP, 1.0);
            __syncthreads();
            diag_block_factor(e, i, sP, sX, &gjb, &s_fail, fail);
        } else {
            double *Kii = tile_ptr(M, T, i, i);
            acc_rsub_global(acc, Kii);
            acc_to_global(acc, Kii);
        }
    }
}

// ------------------------------------------------------------------ potrf: one left-looking step
// grid (evaluations, 1 + rows below).  CTA y == 0 factors diagonal block k (already complete:
// every earlier panel CTA folded its tile into it) while the CTAs y >= 1 run the k-tile
// products of their panel tiles; they then wait on flags[k] (set by the diagonal CTA of the same
// evaluation, which was dispatched before them), finish L_ik = P X_kk^T, push it into the
// right-hand sides and fold it into their own diagonal block.  One launch per block column,
// with the latency-bound diagonal factorisation hidden behind the tensor-core work.
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_potrf_step(const EvalDesc *__restrict__ descs, int k, int *__restrict__ fail)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    __shared__ __align__(16) GjBufs gjb;
    __shared__ double red[2 * MEDGP_NB];
    __shared__ int s_fail;
    const EvalDesc &e = descs[blockIdx.x];  // x = evaluation: ALL diagonal CTAs (y == 0) are dispatched first
    const int T = e.T;
    double *M = e.M;
    double *sP = smem, *sX = smem + kTileElems;
    if (blockIdx.y == 0) {
        if (k >= T) return;
        if (threadIdx.x == 0) s_fail = 0;
        __syncthreads();
        diag_block_factor(e, k, sP, sX, &gjb, &s_fail, fail, false);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) flag_release(e.flags + k);
        return;
    }
    const int i = k + blockIdx.y;
    if (i >= T) return;
    double *Tik = tile_ptr(M, T, i, k);
    const double *Xk = e.dinv + (size_t)k * kTileElems;
    prefetch_tile_l2(Tik);
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    gemm_nt_tiles(acc, k,
                  [&](int l, const double *&A, const double *&B) {
                      A = tile_ptr(M, T, i, l);
                      B = tile_ptr(M, T, k, l);
                  },
                  smem, &bars);
    acc_rsub_global(acc, Tik);  // P = K_ik - C (does not depend on the diagonal block)
    __syncthreads();            // every warp is done with the pipeline buffers
    acc_to_smem(acc, sP, 1.0);
    if (threadIdx.x == 0)
        while (flag_acquire(e.flags + k) == 0) __nanosleep(64);
    __syncthreads();
    tile_bulk_g2s(sX, Xk, &bars);
    tile_bulk_wait(&bars);
    __syncthreads();
    gemm2_smem(acc, sP, sX);
    acc_to_global(acc, Tik);
    acc_matvec_rhs(acc, e, e.rhs + k * MEDGP_NB, e.rhs + i * MEDGP_NB, red);
    // fold into the own diagonal block: K_ii -= L_ik L_ik^T
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    acc_to_smem(acc, sP, 1.0);
    __syncthreads();
    acc_zero(acc);
    mma_panels(acc, sP, sP, MEDGP_NB / 4, warp & 1, warp >> 1, lane);
    double *Kii = tile_ptr(M, T, i, i);
    acc_rsub_global(acc, Kii);
    acc_to_global(acc, Kii);
}

// ------------------------------------------------------------------ trtri: block row i of L^-1
// grid (j < i, evaluations): U_ji = -(sum_{l=j}^{i-1} U_jl L_il^T) X_ii^T, U_jj = X_jj^T
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_trtri_row(const EvalDesc *__restrict__ descs, int i, int right_looking)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    const int j = blockIdx.x;
    if (i >= e.T || j >= i) return;
    const int T = e.T;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    const double *XTj = e.dinvT + (size_t)j * kTileElems;
    if (right_looking) {
        // the sum was accumulated into tile (j, i) by k_trtri_update
        acc_rsub_global(acc, tile_ptr(M, T, j, i));
    } else {
        gemm_nt_tiles(acc, i - j,
                      [&](int l0, const double *&A, const double *&B) {
                          const int l = j + l0;
                          A = (l0 == 0) ? XTj : tile_ptr(M, T, j, l);
                          B = tile_ptr(M, T, i, l);
                      },
                      smem, &bars);
    }
    __syncthreads();
    double *sP = smem, *sX = smem + kTileElems;
    tile_bulk_g2s(sX, e.dinv + (size_t)i * kTileElems, &bars);
    acc_to_smem(acc, sP, -1.0);
    tile_bulk_wait(&bars);
    __syncthreads();
    gemm2_smem(acc, sP, sX);
    acc_to_global(acc, tile_ptr(M, T, j, i));
}

// ------------------------------------------------------------------ right-looking variants
// For few, large matrices the left-looking panel has too few CTAs per launch ((T-k-1) per
// matrix, each k tiles deep).  The right-looking form exposes (T-k)^2/2 independent one-tile
// products per step instead, at the price of re-reading the trailing tiles T times:
//   potrf:  K_ij -= L_ik L_jk^T            for k < j <= i      (k_syrk_update, after step k)
//   trtri:  Acc_ji (+)= U_jk L_ik^T        for j <= k < i      (k_trtri_update), then
//           U_j,k+1 = -Acc_j,k+1 X_k+1^T                       (k_trtri_row, right_looking = 1)
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_syrk_update(const EvalDesc *__restrict__ descs, int k)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    int a, b;
    tri_index(blockIdx.x, a, b);
    const int i = k + 1 + a, j = k + 1 + b, T = e.T;
    if (i >= T) return;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    gemm_nt_tiles(acc, 1,
                  [&](int, const double *&A, const double *&B) {
                      A = tile_ptr(M, T, i, k);
                      B = tile_ptr(M, T, j, k);
                  },
                  smem, &bars);
    double *Cij = tile_ptr(M, T, i, j);
    acc_rsub_global(acc, Cij);
    acc_to_global(acc, Cij);
}

__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_trtri_update(const EvalDesc *__restrict__ descs, int k)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    const int T = e.T;
    const int j = blockIdx.x % (k + 1), i = k + 1 + blockIdx.x / (k + 1);
    if (i >= T) return;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    const double *XTk = e.dinvT + (size_t)k * kTileElems;
    gemm_nt_tiles(acc, 1,
                  [&](int, const double *&A, const double *&B) {
                      A = (j == k) ? XTk : tile_ptr(M, T, j, k);  // U_jk (U_kk = X_kk^T)
                      B = tile_ptr(M, T, i, k);                   // L_ik
                  },
                  smem, &bars);
    double *Aji = tile_ptr(M, T, j, i);
    if (j != k) {  // first touch (j == k) initialises the accumulator tile
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int wm = warp & 1, wn = warp >> 1, r = lane >> 2, kq = lane & 3;
#pragma unroll
        for (int x = 0; x < 4; x++)
#pragma unroll
            for (int y = 0; y < 4; y++) {
                const int row = wm * 32 + 8 * x + r, col = wn * 32 + 8 * y + 2 * kq;
                acc[x][y][0] += Aji[col * MEDGP_SLD + row];
                acc[x][y][1] += Aji[(col + 1) * MEDGP_SLD + row];
            }
    }
    acc_to_global(acc, Aji);
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
    const int T = e.T;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    const double *XTi = e.dinvT + (size_t)i * kTileElems;
    gemm_nt_tiles(acc, T - i,
                  [&](int l0, const double *&A, const double *&B) {
                      const int l = i + l0;
                      A = (l0 == 0) ? XTi : tile_ptr(M, T, i, l);
                      B = (l == j) ? XTi : tile_ptr(M, T, j, l);  // l == j only when i == j == l
                  },
                  smem, &bars);
    acc_to_global(acc, tile_ptr(M, T, i, j));
}

// ------------------------------------------------------------------ NLML
// One CTA per evaluation, after the factorisation (which carried the forward solve along):
//   nlml = 1/2 z^T z + sum log L_ii + n log(2 PI)/2     (c_inference_exact.cpp:118-120,146-152)
__global__ void __launch_bounds__(256)
k_solve(const EvalDesc *__restrict__ descs, ModelDims md, double *__restrict__ out_nlml,
        int *__restrict__ out_status, const int *__restrict__ fail)
{
    __shared__ double scratch[64];
    const EvalDesc &e = descs[blockIdx.x];
    const int ld = e.npad, tid = threadIdx.x;
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
    const double *XT = e.dinvT + (size_t)j * kTileElems;
    for (int c = sl * 16; c < sl * 16 + 16; c++)
        if (c >= r) s += XT[c * MEDGP_SLD + r] * z[j * MEDGP_NB + c];
    // strictly upper tiles (j, l), l > j : columns split over the 4 slices
    for (int c = (j + 1) * MEDGP_NB + sl; c < ld; c += 4)
        s += e.M[tile_off(e.T, j, c >> 6) + (c & 63) * MEDGP_SLD + r] * z[c];
    sp[sl * MEDGP_NB + r] = s;
    __syncthreads();
    if (tid < MEDGP_NB)
        e.alpha[j * MEDGP_NB + tid] = sp[tid] + sp[MEDGP_NB + tid] + sp[2 * MEDGP_NB + tid] + sp[3 * MEDGP_NB + tid];
}
