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
// sP[c][m] (column-major), sX[c][n] = X(n,c), X LOWER triangular: X(n, c) = 0 for c > n, so the
// 8-column sub-tile starting at column n0 only needs the k-steps (of 4) below (n0 + 8) / 4 --
// 288 DMMAs per tile instead of 512 (the warp rotation of gemm_warp() evens the two column
// halves out over the SM's four pipes).  RAGGED: the tile touches the end of the matrix -- only
// mx / ny sub-tiles of this warp hold rows / columns below n and P is zero from column 4 k4max on.
template <int WN, bool RAGGED>
__device__ __forceinline__ void gemm2_tri_half(double (&acc)[4][4][2], const double *sP, const double *sX,
                                               int wm, int lane, int mx, int ny, int k4max)
{
    const int r = lane >> 2, kq = lane & 3;
    const double *pa = sP + kq * MEDGP_SLD + wm * 32 + r;
    const double *pb = sX + kq * MEDGP_SLD + WN * 32 + r;
#pragma unroll
    for (int kk = 0; kk < 8 * WN + 8; kk++) {
        if (RAGGED && kk >= k4max) break;
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            a[u] = pa[kk * 4 * MEDGP_SLD + 8 * u];
            if (kk < 8 * WN + 2 * u + 2) b[u] = pb[kk * 4 * MEDGP_SLD + 8 * u];
        }
#pragma unroll
        for (int x = 0; x < 4; x++) {
            if (RAGGED && x >= mx) continue;
#pragma unroll
            for (int y = 0; y < 4; y++) {
                if (kk >= 8 * WN + 2 * y + 2) continue;
                if (RAGGED && y >= ny) continue;
                dmma884(acc[x][y][0], acc[x][y][1], a[x], b[y]);
            }
        }
    }
}

// m_valid / n_valid: rows / columns of the result that can be non-zero (64 away from the ragged
// end); with n_valid < 64 the columns of sP from n_valid on are zero as well (trtri, last row)
__device__ __forceinline__ void gemm2_smem(double (&acc)[4][4][2], const double *sP, const double *sX,
                                           int m_valid = MEDGP_NB, int n_valid = MEDGP_NB)
{
    const int warp = gemm_warp(), lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = warp >> 1;
    acc_zero(acc);
    if (m_valid < MEDGP_NB || n_valid < MEDGP_NB) {
        const int mx = edge_subtiles(m_valid, wm), ny = edge_subtiles(n_valid, wn);
        const int k4max = n_valid < MEDGP_NB ? (n_valid + 3) >> 2 : MEDGP_NB / 4;
        if (wn == 0) gemm2_tri_half<0, true>(acc, sP, sX, wm, lane, mx, ny, k4max);
        else gemm2_tri_half<1, true>(acc, sP, sX, wm, lane, mx, ny, k4max);
    } else {
        if (wn == 0) gemm2_tri_half<0, false>(acc, sP, sX, wm, lane, 4, 4, MEDGP_NB / 4);
        else gemm2_tri_half<1, false>(acc, sP, sX, wm, lane, 4, 4, MEDGP_NB / 4);
    }
}

// symmetric product acc = sP sP^T of which only the lower part is used (the fold of a fresh panel
// tile into its diagonal block): the upper-right quadrant is skipped; valid = rows of sP below n
__device__ __forceinline__ void syrk_smem(double (&acc)[4][4][2], const double *sP, int valid = MEDGP_NB)
{
    const int warp = gemm_warp(), lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = warp >> 1;
    acc_zero(acc);
    if (wm == 0 && wn == 1) return;
    if (valid < MEDGP_NB)
        mma_panels_edge(acc, sP, sP, MEDGP_NB / 4, wm, wn, lane, edge_subtiles(valid, wm), edge_subtiles(valid, wn));
    else
        mma_panels(acc, sP, sP, MEDGP_NB / 4, wm, wn, lane);
}

// rows of the last block row of an evaluation that lie below n (64 for every other block row)
__device__ __forceinline__ int rows_valid(const EvalDesc &e, int ti)
{
    return ti == e.T - 1 ? e.n - MEDGP_NB * (e.T - 1) : MEDGP_NB;
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
                                                double *red /*3*64*/, bool have_v0 = false, double v0 = 0.0)
{
    // have_v0: thread tid < 64 already holds vec_base[tid] of right-hand side 0 (fetched early)
    const int tid = threadIdx.x, r = tid & 63, half = tid >> 6, ld = e.npad;
    double *vs = red + 2 * MEDGP_NB;
    for (int q = 0; q < e.nrhs; q++) {
        if (tid < MEDGP_NB) vs[tid] = (q == 0 && have_v0) ? v0 : vec_base[(size_t)q * ld + tid];
        __syncthreads();
        double s = 0.0;
        const int cmax = lower_only ? r : MEDGP_NB - 1;
#pragma unroll 4
        for (int c = half; c <= cmax; c += 2) s += sTile[c * MEDGP_SLD + r] * vs[c];
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
    const int warp = gemm_warp(), lane = threadIdx.x & 31, ld = e.npad;
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
// Cholesky factor L and triangular inverse X = L^-1 of one 64x64 block held in shared memory,
// 128 threads, blocked by 16 columns so that the only serial chain is the 64 pivots:
//   for J = 0..3 (column block c0 = 16 J):
//     (a) warp 0 factors the 16x16 diagonal block in registers (lane = row), exchanging the
//         pivot column by shuffles; the reciprocal square root of the NEXT pivot is started
//         before the rest of the rank-1 update is issued, so per pivot the chain is
//         shuffle -> multiply -> fma -> rsqrt.
//     (b) warps 0-1 (lane = row) solve the rows below against the block, L_IJ = S_IJ L_JJ^-T,
//         by right-looking substitution in registers, while warp 2 inverts the diagonal block
//         (lane = column).
//     (c) all warps apply the rank-16 update to the trailing lower part with DMMA.
//   then the off-diagonal blocks X_IJ = -X_II sum_K L_IK X_KJ, level by level, with DMMA.
// LAPACK semantics: a non-positive (or NaN) pivot raises *s_fail (potrf info > 0).
#define MEDGP_DIAG_THREADS 128
// cycle stamps of the diagonal-block routine for tools/diag_phase.cu (no-op in the library)
#ifndef MEDGP_PHASE
#define MEDGP_PHASE(id)
#endif

// scratch of the diagonal-block routines: d[0] holds 1/L_cc during the factorisation and is the
// reduction buffer of the forward-solve matvec afterwards
struct GjBufs {
    double d[3][MEDGP_NB];
};

// 1/sqrt(d): hardware approximation (about 22 bits) refined by one third-order step,
// y1 = y0 (1 + e/2 + 3 e^2/8), e = 1 - d y0^2: error O(e^3), below the rounding of the result.
__device__ __forceinline__ double rsqrt_fast(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
}

// (a) 16x16 diagonal block at (c0, c0): in-place Cholesky (lower part), 1/L_cc -> s_rs[c0 + c].
// One warp, lane = row (lanes 16..31 mirror lanes 0..15 and store nothing).  The pivot's
// reciprocal square root travels by shuffle (it is on the serial chain); the pivot column is
// stored to the tile itself -- its final place -- and read back as broadcasts.
__device__ __forceinline__ void chol16_warp(double *sA, double *s_rs, int c0, int lane, int *s_fail)
{
    const int r = lane & 15;
    double a[16];
#pragma unroll
    for (int c = 0; c < 16; c++) a[c] = (c <= r) ? sA[(c0 + c) * MEDGP_SLD + c0 + r] : 0.0;
    __syncwarp();  // the mirror lanes have read the block before its columns are overwritten below
    double rs = rsqrt_fast(a[0]), rs_mine = rs;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        bad = bad || (r == j && !(a[j] > 0.0));
        rs_mine = (r == j) ? rs : rs_mine;
        const double rsj = __shfl_sync(0xffffffffu, rs, j);
        const double l = a[j] * rsj;  // L_rj for r >= j (r == j: sqrt of the pivot)
        double *col = sA + (c0 + j) * MEDGP_SLD + c0;
        if (lane < 16 && r >= j) col[r] = l;
        if (j < 15) {
            rs = rsqrt_fast(fma(-l, l, a[j + 1]));  // next pivot: meaningful in lane j + 1
            __syncwarp();
            if (j & 1) {  // rows j+1 .. 15 of column j; 16-byte aligned pairs start at an even row
#pragma unroll
                for (int c = j + 1; c < 16; c += 2) {
                    const double2 lc = *reinterpret_cast<const double2 *>(col + c);
                    a[c] = fma(-l, lc.x, a[c]);
                    a[c + 1] = fma(-l, lc.y, a[c + 1]);
                }
            } else {
                a[j + 1] = fma(-l, col[j + 1], a[j + 1]);
#pragma unroll
                for (int c = j + 2; c < 16; c += 2) {
                    const double2 lc = *reinterpret_cast<const double2 *>(col + c);
                    a[c] = fma(-l, lc.x, a[c]);
                    a[c + 1] = fma(-l, lc.y, a[c + 1]);
                }
            }
        }
    }
    if (lane < 16) {
        if (bad) *s_fail = 1;
        s_rs[c0 + r] = rs_mine;
    }
}

// (b) one row below the diagonal block: l_rj = (s_rj - sum_{c<j} l_rc L_jc) / L_jj, right-looking
__device__ __forceinline__ void panel16_row(double *sA, const double *s_rs, int c0, int row)
{
    double a[16];
#pragma unroll
    for (int c = 0; c < 16; c++) a[c] = sA[(c0 + c) * MEDGP_SLD + row];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const double l = a[j] * s_rs[c0 + j];
        a[j] = l;
#pragma unroll
        for (int c = j + 1; c < 16; c++) a[c] = fma(-l, sA[(c0 + j) * MEDGP_SLD + c0 + c], a[c]);
    }
#pragma unroll
    for (int c = 0; c < 16; c++) sA[(c0 + c) * MEDGP_SLD + row] = a[c];
}

// (b') inverse of the 16x16 diagonal block of L into sX (full block, zeros above the diagonal).
// One warp, lane = column c: x_cc = 1/L_cc, x_kc = -(sum_{m<k} L_km x_mc) / L_kk.
__device__ __forceinline__ void trinv16_warp(const double *sA, double *sX, const double *s_rs, int c0, int lane)
{
    const int c = lane & 15;
    double acc[16], x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const double rk = s_rs[c0 + k];
        const double xk = (k == c) ? rk : ((k > c) ? -rk * acc[k] : 0.0);
        x[k] = xk;
#pragma unroll
        for (int r = k + 1; r < 16; r++) acc[r] = fma(sA[(c0 + k) * MEDGP_SLD + c0 + r], xk, acc[r]);
    }
    if (lane < 16) {
#pragma unroll
        for (int r = 0; r < 16; r++) sX[(c0 + c) * MEDGP_SLD + c0 + r] = x[r];
    }
}

// (c) trailing update S -= P P^T, P = the 16 columns at c0, rows/columns >= c0 + 16, lower
// 8x8 tiles dealt round-robin to the 4 warps
__device__ __forceinline__ void trail16_update(double *sA, int c0, int warp, int lane)
{
    const int c1 = c0 + 16, m = (MEDGP_NB - c1) / 8;
    const int lr = lane >> 2, lk = lane & 3;
    int cnt = 0;
    for (int ti = 0; ti < m; ti++)
        for (int tj = 0; tj <= ti; tj++, cnt++) {
            if ((cnt & 3) != warp) continue;
            const int row0 = c1 + 8 * ti, col0 = c1 + 8 * tj;
            double *pc = sA + (col0 + 2 * lk) * MEDGP_SLD + row0 + lr;
            double v0 = pc[0], v1 = pc[MEDGP_SLD];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const double *pk = sA + (c0 + 4 * kk + lk) * MEDGP_SLD;
                dmma884(v0, v1, -pk[row0 + lr], pk[col0 + lr]);
            }
            pc[0] = v0;
            pc[MEDGP_SLD] = v1;
        }
}

// sA: the SPD block (lower part; element (r, c) at c*SLD + r) -> L in place.  sX -> X = L^-1
// (full tile, zeros above the diagonal).  Call with all 128 threads after a barrier that makes
// sA visible; returns after a barrier.
__device__ __forceinline__ void potf2_inv_blocked(double *sA, double *sX, double *s_rs, int *s_fail)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lr = lane >> 2, lk = lane & 3;
    for (int i = tid; i < kTileElems; i += MEDGP_DIAG_THREADS) sX[i] = 0.0;
    MEDGP_PHASE(2)
#pragma unroll 1
    for (int J = 0; J < 4; J++) {
        const int c0 = 16 * J;
        if (warp == 0) chol16_warp(sA, s_rs, c0, lane, s_fail);
        __syncthreads();
        MEDGP_PHASE(3 + 3 * J)
        if (J < 3) {
            if (warp < 2) {
                if (tid >= c0 + 16) panel16_row(sA, s_rs, c0, tid);
            } else if (warp == 2) {
                trinv16_warp(sA, sX, s_rs, c0, lane);
            }
            __syncthreads();
            MEDGP_PHASE(4 + 3 * J)
            trail16_update(sA, c0, warp, lane);
            __syncthreads();
            MEDGP_PHASE(5 + 3 * J)
        } else if (warp == 2) {
            trinv16_warp(sA, sX, s_rs, c0, lane);
        }
    }
    MEDGP_PHASE(13)
    // off-diagonal blocks of X, level d = I - J
#pragma unroll 1
    for (int d = 1; d < 4; d++) {
        __syncthreads();
        if (warp < 4 - d) {
            const int J = warp, I = J + d;
            double t[2][2][2];
#pragma unroll
            for (int u = 0; u < 8; u++) (&t[0][0][0])[u] = 0.0;
            for (int K = J; K < I; K++)  // T = sum_K L_IK X_KJ
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const double *pa = sA + (16 * K + 4 * kk + lk) * MEDGP_SLD + 16 * I + lr;
                    const double *pb = sX + (16 * J + lr) * MEDGP_SLD + 16 * K + 4 * kk + lk;
                    const double a0 = pa[0], a1 = pa[8], b0 = pb[0], b1 = pb[8 * MEDGP_SLD];
                    dmma884(t[0][0][0], t[0][0][1], a0, b0);
                    dmma884(t[0][1][0], t[0][1][1], a0, b1);
                    dmma884(t[1][0][0], t[1][0][1], a1, b0);
                    dmma884(t[1][1][0], t[1][1][1], a1, b1);
                }
            double *px = sX + (16 * J + 2 * lk) * MEDGP_SLD + 16 * I + lr;  // block (I, J), this lane's C slots
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    px[8 * nt * MEDGP_SLD + 8 * mt] = t[mt][nt][0];
                    px[(8 * nt + 1) * MEDGP_SLD + 8 * mt] = t[mt][nt][1];
                }
            __syncwarp();
            double x[2][2][2];
#pragma unroll
            for (int u = 0; u < 8; u++) (&x[0][0][0])[u] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {  // X_IJ = -X_II T
                const double *pa = sX + (16 * I + 4 * kk + lk) * MEDGP_SLD + 16 * I + lr;
                const double *pb = sX + (16 * J + lr) * MEDGP_SLD + 16 * I + 4 * kk + lk;
                const double a0 = -pa[0], a1 = -pa[8], b0 = pb[0], b1 = pb[8 * MEDGP_SLD];
                dmma884(x[0][0][0], x[0][0][1], a0, b0);
                dmma884(x[0][1][0], x[0][1][1], a0, b1);
                dmma884(x[1][0][0], x[1][0][1], a1, b0);
                dmma884(x[1][1][0], x[1][1][1], a1, b1);
            }
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    px[8 * nt * MEDGP_SLD + 8 * mt] = x[mt][nt][0];
                    px[(8 * nt + 1) * MEDGP_SLD + 8 * mt] = x[mt][nt][1];
                }
        }
    }
    __syncthreads();
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
    const double *Kkk = tile_ptr(e.M, T, k, k);
    MEDGP_PHASE(0)
    // all global loads of this block are issued up front: one memory round trip
    double2 kv[16];
#pragma unroll
    for (int u = 0; u < 16; u++) {
        const int idx = tid + MEDGP_DIAG_THREADS * u, c = idx >> 5, rp = idx & 31;
        kv[u] = *reinterpret_cast<const double2 *>(Kkk + c * MEDGP_SLD + 2 * rp);
    }
    const double v0 = (tid < MEDGP_NB) ? e.rhs[k * MEDGP_NB + tid] : 0.0;
#pragma unroll
    for (int u = 0; u < 16; u++) {
        const int idx = tid + MEDGP_DIAG_THREADS * u, c = idx >> 5, rp = idx & 31, o = c * MEDGP_SLD + 2 * rp;
        double2 v = kv[u];
        if (have_product) {
            const double2 pr = *reinterpret_cast<const double2 *>(sD + o);
            v.x -= pr.x;
            v.y -= pr.y;
        }
        if (2 * rp < c) v.x = 0.0;
        if (2 * rp + 1 < c) v.y = 0.0;
        *reinterpret_cast<double2 *>(sL + o) = v;
    }
    __syncthreads();  // sD is dead from here on: it receives X
    MEDGP_PHASE(1)
    potf2_inv_blocked(sL, sD, gjb->d[0], s_fail);
    MEDGP_PHASE(14)
    // forward solve, block k: z_k = X_kk rhs_k (in place)
    tile_matvec_rhs(e, sD, e.rhs + k * MEDGP_NB, e.rhs + k * MEDGP_NB, true, false, gjb->d[0], true, v0);
    MEDGP_PHASE(15)
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
    MEDGP_PHASE(16)
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
    if (k >= e.T || e.skip) return;
    const int T = e.T, tid = threadIdx.x;
    if (depth > 0) prefetch_tile_l2(tile_ptr(e.M, T, k, k));  // wanted right after the products
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
                      smem, &bars, NoStageFn(),
                      // only the lower part of the symmetric product is used
                      [](int, int wm, int wn) { return wm == 0 && wn == 1; },
                      TileEdge{rows_valid(e, k), rows_valid(e, k), MEDGP_NB});
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
    if (i >= e.T || e.skip) return;
    const int T = e.T;
    double *M = e.M;
    double *Tik = tile_ptr(M, T, i, k);
    const double *Xk = e.dinv + (size_t)k * kTileElems;
    __shared__ double red[2 * MEDGP_NB];
    prefetch_tile_l2(Tik);  // epilogue operands: start them towards L2 now
    prefetch_tile_l2(Xk);
    gemm_bars_init(&bars);
    const int mv = rows_valid(e, i);  // the last block row is zero from row mv on
    double acc[4][4][2];
    acc_zero(acc);
    gemm_nt_tiles(acc, depth,
                  [&](int l, const double *&A, const double *&B) {
                      A = tile_ptr(M, T, i, l);
                      B = tile_ptr(M, T, k, l);
                  },
                  smem, &bars, NoStageFn(), NoSkipFn(), TileEdge{mv, MEDGP_NB, MEDGP_NB});
    __syncthreads();  // every warp is done with the pipeline buffers
    double *sP = smem, *sX = smem + kTileElems;
    tile_bulk_g2s(sX, Xk, &bars);  // X_kk arrives while P = K_ik - C is formed
    acc_rsub_global(acc, Tik);
    acc_to_smem(acc, sP, 1.0);
    tile_bulk_wait(&bars);
    __syncthreads();
    gemm2_smem(acc, sP, sX, mv);
    acc_to_global(acc, Tik);
    // forward solve: push the fresh tile into the right-hand sides of block row i
    acc_matvec_rhs(acc, e, e.rhs + k * MEDGP_NB, e.rhs + i * MEDGP_NB, red);
    if (fold_diag) {
        // apply this tile to its diagonal block right away, K_ii -= L_ik L_ik^T, so that no
        // diagonal kernel has a k-tile product to do.  With fold_diag == 2 the CTA of row k+1,
        // whose diagonal block is now complete, factors it on the spot: the next step's
        // diagonal kernel disappears and its latency hides behind the other panel CTAs.
        acc_to_smem(acc, sP, 1.0);  // (acc_matvec_rhs ended with a block barrier: GEMM2 is done with sP)
        __syncthreads();
        syrk_smem(acc, sP, mv);
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
// One launch per block column for chunks of few matrices.  A CTA's ROLE is not its block index
// but a ticket drawn from a per-launch counter when it starts running: tickets 0 .. ndiag-1
// factor diagonal block k of evaluation `ticket` (already complete: every earlier panel CTA
// folded its tile into it); later tickets are panel tiles (evaluation, row), which run their
// k-tile products, then wait on flags[k] of their evaluation, finish L_ik = P X_kk^T, push it
// into the right-hand sides and fold it into their own diagonal block.  Because a ticket is only
// ever held by a CTA that is already executing, every flag a panel CTA can wait for belongs to a
// diagonal role that is running or finished -- forward progress does not depend on the order in
// which the hardware dispatches blocks, whatever the grid size.  The latency-bound diagonal
// factorisation hides behind the tensor-core work of the panel roles.
// *ticket must be 0 at launch (k_prep zeroes the sub-chunk's counters).
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_potrf_step(const EvalDesc *__restrict__ descs, int k, int *__restrict__ fail, int *__restrict__ ticket)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    __shared__ __align__(16) GjBufs gjb;
    __shared__ double red[2 * MEDGP_NB];
    __shared__ int s_fail, s_role;
    if (threadIdx.x == 0) s_role = atomicAdd(ticket, 1);
    __syncthreads();
    const int ndiag = gridDim.x, role = s_role;
    const int ev = role < ndiag ? role : (role - ndiag) % ndiag;
    const int row = role < ndiag ? 0 : 1 + (role - ndiag) / ndiag;
    const EvalDesc &e = descs[ev];
    if (e.skip) return;
    const int T = e.T;
    double *M = e.M;
    double *sP = smem, *sX = smem + kTileElems;
    if (row == 0) {
        if (k >= T) return;
        if (threadIdx.x == 0) s_fail = 0;
        __syncthreads();
        diag_block_factor(e, k, sP, sX, &gjb, &s_fail, fail, false);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) flag_release(e.flags + k);
        return;
    }
    const int i = k + row;
    if (i >= T) return;
    double *Tik = tile_ptr(M, T, i, k);
    const double *Xk = e.dinv + (size_t)k * kTileElems;
    prefetch_tile_l2(Tik);
    gemm_bars_init(&bars);
    const int mv = rows_valid(e, i);
    double acc[4][4][2];
    acc_zero(acc);
    gemm_nt_tiles(acc, k,
                  [&](int l, const double *&A, const double *&B) {
                      A = tile_ptr(M, T, i, l);
                      B = tile_ptr(M, T, k, l);
                  },
                  smem, &bars, NoStageFn(), NoSkipFn(), TileEdge{mv, MEDGP_NB, MEDGP_NB});
    acc_rsub_global(acc, Tik);  // P = K_ik - C (does not depend on the diagonal block)
    __syncthreads();            // every warp is done with the pipeline buffers
    acc_to_smem(acc, sP, 1.0);
    if (threadIdx.x == 0)
        while (flag_acquire(e.flags + k) == 0) __nanosleep(64);
    __syncthreads();
    tile_bulk_g2s(sX, Xk, &bars);
    tile_bulk_wait(&bars);
    __syncthreads();
    gemm2_smem(acc, sP, sX, mv);
    acc_to_global(acc, Tik);
    acc_matvec_rhs(acc, e, e.rhs + k * MEDGP_NB, e.rhs + i * MEDGP_NB, red);
    // fold into the own diagonal block: K_ii -= L_ik L_ik^T
    acc_to_smem(acc, sP, 1.0);
    __syncthreads();
    syrk_smem(acc, sP, mv);
    double *Kii = tile_ptr(M, T, i, i);
    acc_rsub_global(acc, Kii);
    acc_to_global(acc, Kii);
}

// ------------------------------------------------------------------ device-side jitter loop
// Last node of a chunk's launch sequence, which is the body of a CUDA-graph WHILE node.  For
// every evaluation of the chunk: a failed factorisation with jitter left gets one more noise
// addition (K_ii += sigma^2 again, inference/c_inference_exact.cpp:99-108) and runs again in the
// next pass; everything else is final and is skipped from now on.  The loop ends when no
// evaluation is left.  One CTA.
__global__ void __launch_bounds__(1024)
k_retry_decide(EvalDesc *__restrict__ descs, int count, int *__restrict__ fail, int max_jitter,
               cudaGraphConditionalHandle handle)
{
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    int any = 0;
    for (int b = threadIdx.x; b < count; b += blockDim.x) {
        EvalDesc &e = descs[b];
        if (e.skip) continue;
#if defined(MEDGP_X_NODMMA) || defined(MEDGP_X_NOLOAD)
        fail[e.out_index] = 0;  // timing experiments produce garbage: one pass only
#endif
        if (fail[e.out_index] != 0 && e.jitter < max_jitter) {
            e.jitter++;
            fail[e.out_index] = 0;
            any = 1;
        } else {
            e.skip = 1;
        }
    }
    if (any) atomicOr(&s_any, 1);
    __syncthreads();
    if (threadIdx.x == 0) cudaGraphSetConditional(handle, (unsigned)s_any);
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
    if (i >= e.T || j >= i || e.skip) return;
    const int T = e.T, nv = rows_valid(e, i);
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
                      smem, &bars, NoStageFn(),
                      // U_jj = X_jj^T is upper triangular: rows 32..63 vanish in k-panels 0 and 1
                      [](int ch, int wm, int) { return ch < 2 && wm == 1; },
                      // block row i of L ends at row nv: columns nv.. of U_ji are zero
                      TileEdge{MEDGP_NB, nv, MEDGP_NB});
    }
    __syncthreads();
    double *sP = smem, *sX = smem + kTileElems;
    tile_bulk_g2s(sX, e.dinv + (size_t)i * kTileElems, &bars);
    acc_to_smem(acc, sP, -1.0);
    tile_bulk_wait(&bars);
    __syncthreads();
    gemm2_smem(acc, sP, sX, MEDGP_NB, nv);
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
k_syrk_update(const EvalDesc *__restrict__ descs, int k, int part)
{
    // part 0: every lower tile of the trailing matrix; 1: its first block column only (the tiles
    // the next step's diagonal and panel kernels need: the critical path of the look-ahead
    // schedule); 2: everything but the first block column (runs beside the next step)
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    int a, b;
    if (part == 1) {
        a = blockIdx.x; b = 0;
    } else {
        tri_index(blockIdx.x, a, b);
        if (part == 2) { a += 1; b += 1; }
    }
    const int i = k + 1 + a, j = k + 1 + b, T = e.T;
    if (i >= T || e.skip) return;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    const bool diag = (i == j);
    gemm_nt_tiles(acc, 1,
                  [&](int, const double *&A, const double *&B) {
                      A = tile_ptr(M, T, i, k);
                      B = tile_ptr(M, T, j, k);
                  },
                  smem, &bars, NoStageFn(),
                  // diagonal tiles: only the lower part of the symmetric update is used
                  [=](int, int wm, int wn) { return diag && wm == 0 && wn == 1; },
                  TileEdge{rows_valid(e, i), rows_valid(e, j), MEDGP_NB});
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
    if (i >= T || e.skip) return;
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
                  smem, &bars, NoStageFn(),
                  // U_kk is upper triangular: rows 32..63 vanish in k-panels 0 and 1
                  [=](int ch, int wm, int) { return j == k && ch < 2 && wm == 1; },
                  TileEdge{MEDGP_NB, rows_valid(e, i), MEDGP_NB});
    double *Aji = tile_ptr(M, T, j, i);
    if (j != k) {  // first touch (j == k) initialises the accumulator tile
        const int warp = gemm_warp(), lane = threadIdx.x & 31;
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
// grid (lower tiles, evaluations): (K^-1)_ij = sum_{l>=i} U_il U_jl^T  -> written over L_ij.
// The CTAs of the diagonal tiles stream the whole block row i of U anyway, so they also form
// alpha_i = (L^-T z)_i = sum_l U_il z_l from the resident panels (no separate pass over U).
#ifndef MEDGP_LAUUM_NST
// k_lauum has no two-tile epilogue, so its shared memory is the pipeline alone: 3 stages (52 KB)
// and 128 registers let 4 CTAs share an SM (16 warps) where the other tile kernels run 3
#define MEDGP_LAUUM_NST 3
#define MEDGP_LAUUM_OCC 4
#endif
constexpr int kLauumSmemBytes = MEDGP_LAUUM_NST * kStageElems * 8;
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, MEDGP_LAUUM_OCC)
k_lauum(const EvalDesc *__restrict__ descs)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    __shared__ double s_al[2 * MEDGP_NB];
    const EvalDesc &e = descs[blockIdx.y];
    int i, j;
    tri_index(blockIdx.x, i, j);
    if (i >= e.T || e.skip) return;
    const int T = e.T;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    const double *XTi = e.dinvT + (size_t)i * kTileElems;
    const bool diag = (i == j);
    const int m = threadIdx.x & 63, kh = (threadIdx.x >> 6) * (MEDGP_KC / 2);
    const double *z = e.rhs + i * MEDGP_NB + kh;
    double asum = 0.0;
    gemm_nt_tiles<MEDGP_LAUUM_NST>(acc, T - i,
                  [&](int l0, const double *&A, const double *&B) {
                      const int l = i + l0;
                      A = (l0 == 0) ? XTi : tile_ptr(M, T, i, l);
                      B = (l == j) ? XTi : tile_ptr(M, T, j, l);  // l == j only when i == j == l
                  },
                  smem, &bars,
                  [&](int ch, const double *stage) {
                      if (!diag) return;
                      const double *zc = z + ch * MEDGP_KC;  // columns (i*64 + 16 ch + kh ..) of block row i
                      const double *pa = stage + kh * MEDGP_SLD + m;
#pragma unroll
                      for (int k = 0; k < MEDGP_KC / 2; k++) asum = fma(pa[k * MEDGP_SLD], __ldg(zc + k), asum);
                  },
                  // the first tile of the row is U_ii = X_ii^T, upper triangular: its rows 32..63 are
                  // zero in columns 0..31, i.e. in k-panels 0 and 1 (also as the B operand when i == j)
                  // Diagonal tiles are symmetric and only their lower part is ever read, so their
                  // upper-right quadrant (rows 0..31, columns 32..63) is not computed at all.
                  [&](int ch, int wm, int wn) {
                      return (diag && wm == 0 && wn == 1) || (ch < 2 && (wm == 1 || (diag && wn == 1)));
                  },
                  // the last tiles of block rows i, j < T-1 of U are zero from column n - 64 (T-1) on
                  TileEdge{MEDGP_NB, MEDGP_NB, i < T - 1 ? rows_valid(e, T - 1) : MEDGP_NB});
    acc_to_global(acc, tile_ptr(M, T, i, j));
    if (diag) {
        s_al[threadIdx.x] = asum;
        __syncthreads();
        if (threadIdx.x < MEDGP_NB) e.alpha[i * MEDGP_NB + threadIdx.x] = s_al[threadIdx.x] + s_al[MEDGP_NB + threadIdx.x];
    }
}

// ------------------------------------------------------------------ alpha without K^-1
// alpha = K^-1 y = L^-T z = U z after the triangular inverse, for callers that want the factor
// itself (medgp_cuda_export_factors) and therefore skip k_lauum, which forms alpha on the gradient
// path.  grid (block rows, evaluations), 128 threads: alpha_i = sum_{l >= i} U_il z_l with
// U_ii = X_ii^T (dinvT) and U_il the strictly-upper tiles.
__global__ void __launch_bounds__(128)
k_alpha(const EvalDesc *__restrict__ descs)
{
    __shared__ double red[2 * MEDGP_NB];
    const EvalDesc &e = descs[blockIdx.y];
    const int i = blockIdx.x, T = e.T;
    if (i >= T || e.skip) return;
    const int r = threadIdx.x & 63, half = threadIdx.x >> 6;
    double s = 0.0;
    for (int l = i; l < T; l++) {
        const double *U = (l == i) ? e.dinvT + (size_t)i * kTileElems : e.M + tile_off(T, i, l);
        const double *z = e.rhs + l * MEDGP_NB;
        for (int c = half; c < MEDGP_NB; c += 2) s = fma(U[c * MEDGP_SLD + r], __ldg(z + c), s);
    }
    red[half * MEDGP_NB + r] = s;
    __syncthreads();
    if (half == 0) e.alpha[i * MEDGP_NB + r] = red[r] + red[MEDGP_NB + r];
}

// ------------------------------------------------------------------ NLML
// One CTA per evaluation, after the factorisation (which carried the forward solve along):
//   nlml = 1/2 z^T z + sum log L_ii + n log(2 PI)/2     (c_inference_exact.cpp:118-120,146-152)
__global__ void __launch_bounds__(256)
k_solve(const EvalDesc *__restrict__ descs, ModelDims md, double *__restrict__ out_nlml,
        int *__restrict__ out_status, int *__restrict__ fail, int force_fail)
{
    __shared__ double scratch[64];
    const EvalDesc &e = descs[blockIdx.x];
    if (e.skip) return;
    const int ld = e.npad, tid = threadIdx.x;
    double v[2] = {0.0, 0.0};
    for (int i = tid; i < ld; i += blockDim.x) v[0] += e.rhs[i] * e.rhs[i];
    for (int k = tid; k < e.T; k += blockDim.x) v[1] += e.blk[k];
    block_reduce_sum<2>(v, scratch);
    if (tid == 0) {
        // force_fail (tests of the jitter path): the first attempts count as failed factorisations
#if defined(MEDGP_X_NODMMA) || defined(MEDGP_X_NOLOAD)
        const bool bad = false;  // timing experiments produce garbage: no retries
        fail[e.out_index] = 0;
#else
        const bool bad = fail[e.out_index] != 0 || e.jitter < force_fail;
#endif
        if (bad) fail[e.out_index] = 1;
        const double nlml = 0.5 * v[0] + v[1] + e.n * log(2.0 * md.pi) / 2.0;
        out_nlml[e.out_index] = bad ? __longlong_as_double(0x7ff8000000000000LL) : nlml;
        out_status[e.out_index] = bad ? -1 : e.jitter;
    }
}


// ------------------------------------------------------------------ stream stagger
// Holds a sub-chunk's stream back for `ns` nanoseconds at the start of a step, so that the
// sub-chunks do not march through the latency-bound phases (diagonal blocks, short row kernels)
// in lock-step but interleave them with the other streams' tensor-core phases.
__global__ void k_delay(unsigned long long ns)
{
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(500);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}
