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
// 288 DMMAs per tile instead of 512.  RAGGED: the tile touches the end of the matrix -- only
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
        if (tid < MEDGP_NB) vs[tid] = (q == 0 && have_v0) ? v0 : __ldcg(vec_base + (size_t)q * ld + tid);
        __syncthreads();
        const int cmax = lower_only ? r : MEDGP_NB - 1;
        double s4[4] = {0.0, 0.0, 0.0, 0.0};  // four independent chains
        int c = half;
        for (; c + 6 <= cmax; c += 8) {
#pragma unroll
            for (int u = 0; u < 4; u++) s4[u] = fma(sTile[(c + 2 * u) * MEDGP_SLD + r], vs[c + 2 * u], s4[u]);
        }
        for (; c <= cmax; c += 2) s4[0] = fma(sTile[c * MEDGP_SLD + r], vs[c], s4[0]);
        const double s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
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
            // through L2: in the dataflow kernels the previous update of this block came from another CTA of the same launch
            double *o = out_base + (size_t)q * ld + threadIdx.x;
            __stcg(o, __ldcg(o) - (red[threadIdx.x] + red[MEDGP_NB + threadIdx.x]));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ potrf: diagonal block k
// Cholesky factor L and triangular inverse X = L^-1 of one 64x64 block held in shared memory,
// 128 threads, blocked by 16 columns so that the only serial chain is the 64 pivots.  For
// J = 0..3 (column block c0 = 16 J):
//   (1) warp 0 factors the 16x16 diagonal block in registers: lanes 0-15 hold its rows, and
//       lanes 16-31 run the forward substitution L X = I for the block's inverse in the SAME
//       instruction stream (with v = -e_c as the start vector both recurrences are
//       v[c] -= w L_cj, w = v[j] / L_jj), so X_JJ costs nothing.  Per pivot the chain is
//       shuffle -> multiply -> fma -> rsqrt; the reciprocal square root of the next pivot is
//       started before the rest of the rank-1 update is issued.
//       Meanwhile warps 1-3 do what is off the critical path: the trailing tiles of the previous
//       column block that the next step does not need yet, and the off-diagonal blocks of X
//       whose inputs are complete (X_10 during J = 2; X_20, X_21 during J = 3).
//   (2) the rows below by substitution against L_JJ, one thread per row;
//   (3) all warps: rank-16 update of the NEXT column block only (what step J+1 needs).
// After J = 3 one level is left: X_3J = -X_33 sum_K L_3K X_KJ on warps 0-2 while warp 3 forms the
// block's log-determinant.
// LAPACK semantics: a non-positive (or NaN) pivot raises *s_fail (potrf info > 0).
#define MEDGP_DIAG_THREADS 128
// cycle stamps of the diagonal-block routine for tools/diag_phase.cu (no-op in the library)
#ifndef MEDGP_PHASE
#define MEDGP_PHASE(id)
#endif

// scratch of the diagonal-block routines: the reduction buffer of the forward-solve matvec and
// the block's log-determinant
struct GjBufs {
    double d[3][MEDGP_NB];
    double logdet;
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

// (1) 16x16 diagonal block at (c0, c0): in-place Cholesky (lower part of sA) and its inverse
// (full block of sX, zeros above the diagonal).  One warp: lane r < 16 = row r of the block,
// lane 16 + c = column c of the inverse.  The pivot's reciprocal square root travels by shuffle
// (it is on the serial chain); the pivot column is stored to the tile itself -- its final place
// -- and read back as broadcasts by both halves of the warp.
__device__ __forceinline__ void chol16_inv_warp(double *sA, double *sX, int c0, int lane, int *s_fail)
{
    const int r = lane & 15;
    const bool xl = lane >= 16;
    double v[16];
#pragma unroll
    for (int c = 0; c < 16; c++)
        v[c] = xl ? (c == r ? -1.0 : 0.0) : ((c <= r) ? sA[(c0 + c) * MEDGP_SLD + c0 + r] : 0.0);
    __syncwarp();  // the block has been read before its columns are overwritten below
    double rs = rsqrt_fast(v[0]);
    bool bad = false;
    double *dst = xl ? sX + (c0 + r) * MEDGP_SLD + c0 : sA + c0 * MEDGP_SLD + c0 + r;
    const int stride = xl ? 1 : MEDGP_SLD, sgn = xl ? -1 : 1;
    const double sign = xl ? -1.0 : 1.0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        bad = bad || (!xl && r == j && !(v[j] > 0.0));
        const double rsj = __shfl_sync(0xffffffffu, rs, j);
        const double w = v[j] * rsj;  // rows: L_rj (r == j: sqrt of the pivot); X lanes: -X_jc
        double *col = sA + (c0 + j) * MEDGP_SLD + c0;
        // ONE unconditional store for both halves of the warp (a branch here would split the warp at
        // every pivot): rows write L_rj down column j (zeros above the diagonal, which is what is
        // there already), X lanes write X_jc = -w along their own column (zeros above the diagonal)
        dst[j * stride] = ((r - j) * sgn >= 0) ? sign * w : 0.0;
        if (j < 15) {
            rs = rsqrt_fast(fma(-w, w, v[j + 1]));  // next pivot: meaningful in lane j + 1
            __syncwarp();
            const double p = -w;
            if (j & 1) {  // rows j+1 .. 15 of column j; 16-byte aligned pairs start at an even row
#pragma unroll
                for (int c = j + 1; c < 16; c += 2) {
                    const double2 lc = *reinterpret_cast<const double2 *>(col + c);
                    v[c] = fma(p, lc.x, v[c]);
                    v[c + 1] = fma(p, lc.y, v[c + 1]);
                }
            } else {
                v[j + 1] = fma(p, col[j + 1], v[j + 1]);
#pragma unroll
                for (int c = j + 2; c < 16; c += 2) {
                    const double2 lc = *reinterpret_cast<const double2 *>(col + c);
                    v[c] = fma(p, lc.x, v[c]);
                    v[c + 1] = fma(p, lc.y, v[c + 1]);
                }
            }
        }
    }
    if (!xl && bad) *s_fail = 1;
}

// (2) one row below the diagonal block: l_rj = (s_rj - sum_{c<j} l_rc L_jc) / L_jj by right-looking
// substitution in registers (thread = row).  Substitution, not a product with the explicit inverse
// X_JJ: on ill-conditioned blocks (condition 1e8 and beyond) the product loses an order of
// magnitude of accuracy; 1 / L_jj is the diagonal of X_JJ.
__device__ __forceinline__ void panel16_row(double *sA, const double *sX, int c0, int row)
{
    double a[16];
#pragma unroll
    for (int c = 0; c < 16; c++) a[c] = sA[(c0 + c) * MEDGP_SLD + row];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const double l = a[j] * sX[(c0 + j) * MEDGP_SLD + c0 + j];
        a[j] = l;
#pragma unroll
        for (int c = j + 1; c < 16; c++) a[c] = fma(-l, sA[(c0 + j) * MEDGP_SLD + c0 + c], a[c]);
    }
#pragma unroll
    for (int c = 0; c < 16; c++) sA[(c0 + c) * MEDGP_SLD + row] = a[c];
}

// (3) trailing update S -= P P^T with P = the 16 columns at c0, restricted to the lower 8x8 tiles
// of the column tiles ct_lo .. ct_lo+NCT-1 (counted from column c0 + 16).  Row tiles are dealt
// to nw warps (this one is wi); the tiles of a row are independent DMMA chains.
template <int NCT>
__device__ __forceinline__ void trail16_cols(double *sA, int c0, int ct_lo, int wi, int nw, int lane)
{
    const int c1 = c0 + 16, m = (MEDGP_NB - c1) / 8;
    const int lr = lane >> 2, lk = lane & 3;
    for (int rt = ct_lo + wi; rt < m; rt += nw) {
        const int row0 = c1 + 8 * rt;
        double v[NCT][2], v2[NCT][2];  // two independent chains per tile (even / odd k-steps)
#pragma unroll
        for (int u = 0; u < NCT; u++)
            if (ct_lo + u <= rt) {
                const double *pc = sA + (c1 + 8 * (ct_lo + u) + 2 * lk) * MEDGP_SLD + row0 + lr;
                v[u][0] = pc[0];
                v[u][1] = pc[MEDGP_SLD];
                v2[u][0] = v2[u][1] = 0.0;
            }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const double *pk = sA + (c0 + 4 * kk + lk) * MEDGP_SLD;
            const double a = -pk[row0 + lr];
#pragma unroll
            for (int u = 0; u < NCT; u++)
                if (ct_lo + u <= rt) {
                    const double bb = pk[c1 + 8 * (ct_lo + u) + lr];
                    if (kk & 1) dmma884(v2[u][0], v2[u][1], a, bb);
                    else dmma884(v[u][0], v[u][1], a, bb);
                }
        }
#pragma unroll
        for (int u = 0; u < NCT; u++)
            if (ct_lo + u <= rt) {
                double *pc = sA + (c1 + 8 * (ct_lo + u) + 2 * lk) * MEDGP_SLD + row0 + lr;
                pc[0] = v[u][0] + v2[u][0];
                pc[MEDGP_SLD] = v[u][1] + v2[u][1];
            }
    }
}

// one off-diagonal 16x16 block of the inverse, X_IJ = -X_II sum_{K=J}^{I-1} L_IK X_KJ (I > J): one warp
__device__ __forceinline__ void xblock16(const double *sA, double *sX, int I, int J, int lane)
{
    const int lr = lane >> 2, lk = lane & 3;
    // two accumulator sets per product (even / odd k-steps): a dependent DMMA costs about 50 cycles
    double t[2][2][2][2];
#pragma unroll
    for (int u = 0; u < 16; u++) (&t[0][0][0][0])[u] = 0.0;
    for (int K = J; K < I; K++)  // T = sum_K L_IK X_KJ
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const double *pa = sA + (16 * K + 4 * kk + lk) * MEDGP_SLD + 16 * I + lr;
            const double *pb = sX + (16 * J + lr) * MEDGP_SLD + 16 * K + 4 * kk + lk;
            const double a0 = pa[0], a1 = pa[8], b0 = pb[0], b1 = pb[8 * MEDGP_SLD];
            dmma884(t[kk & 1][0][0][0], t[kk & 1][0][0][1], a0, b0);
            dmma884(t[kk & 1][0][1][0], t[kk & 1][0][1][1], a0, b1);
            dmma884(t[kk & 1][1][0][0], t[kk & 1][1][0][1], a1, b0);
            dmma884(t[kk & 1][1][1][0], t[kk & 1][1][1][1], a1, b1);
        }
    double *px = sX + (16 * J + 2 * lk) * MEDGP_SLD + 16 * I + lr;  // block (I, J), this lane's C slots
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            px[8 * nt * MEDGP_SLD + 8 * mt] = t[0][mt][nt][0] + t[1][mt][nt][0];
            px[(8 * nt + 1) * MEDGP_SLD + 8 * mt] = t[0][mt][nt][1] + t[1][mt][nt][1];
        }
    __syncwarp();
    double x[2][2][2][2];
#pragma unroll
    for (int u = 0; u < 16; u++) (&x[0][0][0][0])[u] = 0.0;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {  // X_IJ = -X_II T
        const double *pa = sX + (16 * I + 4 * kk + lk) * MEDGP_SLD + 16 * I + lr;
        const double *pb = sX + (16 * J + lr) * MEDGP_SLD + 16 * I + 4 * kk + lk;
        const double a0 = -pa[0], a1 = -pa[8], b0 = pb[0], b1 = pb[8 * MEDGP_SLD];
        dmma884(x[kk & 1][0][0][0], x[kk & 1][0][0][1], a0, b0);
        dmma884(x[kk & 1][0][1][0], x[kk & 1][0][1][1], a0, b1);
        dmma884(x[kk & 1][1][0][0], x[kk & 1][1][0][1], a1, b0);
        dmma884(x[kk & 1][1][1][0], x[kk & 1][1][1][1], a1, b1);
    }
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            px[8 * nt * MEDGP_SLD + 8 * mt] = x[0][mt][nt][0] + x[1][mt][nt][0];
            px[(8 * nt + 1) * MEDGP_SLD + 8 * mt] = x[0][mt][nt][1] + x[1][mt][nt][1];
        }
}

// sA: the SPD block (lower part; element (r, c) at c*SLD + r) -> L in place.  sX -> X = L^-1
// (full tile, zeros above the diagonal; no need to clear it first).  *logdet = sum log L_ii.
// Call with all 128 threads after a barrier that makes sA visible; returns after a barrier.
__device__ __forceinline__ void potf2_inv_blocked(double *sA, double *sX, double *logdet, int *s_fail)
{
    // (rotating these warp roles by the block index, so that the factoring warps of the CTAs resident
    // on one SM would not share a sub-partition, was measured: no difference)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll 1
    for (int J = 0; J < 4; J++) {
        const int c0 = 16 * J;
        if (warp == 0) {
            chol16_inv_warp(sA, sX, c0, lane, s_fail);
        } else if (J == 0) {
            // everything of X outside the four diagonal 16x16 blocks starts as zero (the upper
            // blocks stay zero, the lower ones are overwritten; rows 64..67 are pitch padding)
            for (int i = tid - 32; i < kTileElems; i += MEDGP_DIAG_THREADS - 32) {
                const int c = i / MEDGP_SLD, rr = i - c * MEDGP_SLD;
                if (rr >= MEDGP_NB || (rr >> 4) != (c >> 4)) sX[i] = 0.0;
            }
        } else {
            trail16_cols<4>(sA, c0 - 16, 2, warp - 1, 3, lane);  // what step J-1 left for later
            if (J == 2 && warp == 3) xblock16(sA, sX, 1, 0, lane);
            if (J == 3 && warp == 1) xblock16(sA, sX, 2, 1, lane);
            if (J == 3 && warp == 2) xblock16(sA, sX, 2, 0, lane);
        }
        __syncthreads();
        MEDGP_PHASE(3 + 3 * J)
        if (J < 3) {
            if (tid >= c0 + 16 && tid < MEDGP_NB) panel16_row(sA, sX, c0, tid);
            __syncthreads();
            MEDGP_PHASE(4 + 3 * J)
            trail16_cols<2>(sA, c0, 0, warp, 4, lane);
            __syncthreads();
            MEDGP_PHASE(5 + 3 * J)
        }
    }
    MEDGP_PHASE(13)
    if (warp < 3) {
        xblock16(sA, sX, 3, warp, lane);
    } else {
        // (diagonal entries of a factor of a covariance block: far from the range where a product of two overflows)
        double s = log(sA[lane * MEDGP_SLD + lane] * sA[(lane + 32) * MEDGP_SLD + lane + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) *logdet = s;
    }
    __syncthreads();
}

// Factor one diagonal block: D = K_kk - C with the accumulated product C in sD (pitch SLD);
// L_kk = chol(D), X_kk = inv(L_kk).  Writes L_kk (lower part of the tile), dinv[k], dinvT[k],
// blk[k] = sum log diag(L_kk), turns rhs block k into z_k = X_kk rhs_k and raises fail[] when a
// pivot is not positive.  128 threads; sD / sL are the two halves of the dynamic smem ring.
__device__ __forceinline__ void diag_block_factor(const EvalDesc &e, int k, double *sD, double *sL,
                                                  GjBufs *gjb, int *s_fail, int *__restrict__ fail,
                                                  bool have_product = true, int *publish = nullptr, int publish_value = 1)
{
    const int T = e.T, tid = threadIdx.x;
    const double *Kkk = tile_ptr(e.M, T, k, k);
    MEDGP_PHASE(0)
    // all global loads of this block are issued up front: one memory round trip
    double2 kv[16];
#pragma unroll
    for (int u = 0; u < 16; u++) {
        const int idx = tid + MEDGP_DIAG_THREADS * u, c = idx >> 5, rp = idx & 31;
        kv[u] = __ldcg(reinterpret_cast<const double2 *>(Kkk + c * MEDGP_SLD + 2 * rp));  // (through L2: k_potrf_flow's chained roles read a tile another CTA of the launch wrote)
    }
    const double v0 = (tid < MEDGP_NB) ? __ldcg(e.rhs + k * MEDGP_NB + tid) : 0.0;
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
    for (int i = tid; i < MEDGP_NB * (MEDGP_SLD - MEDGP_NB); i += MEDGP_DIAG_THREADS)  // pitch padding: defined bytes for the bulk store
        sL[(i >> 2) * MEDGP_SLD + MEDGP_NB + (i & 3)] = 0.0;
    __syncthreads();  // sD is dead from here on: it receives X
    MEDGP_PHASE(1)
    potf2_inv_blocked(sL, sD, &gjb->logdet, s_fail);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's tile writes -> the bulk stores below
    __syncthreads();
    if (tid == 0) {  // two groups: X_kk (what waiting panel roles need) and L_kk
        bulk_s2g(e.dinv + (size_t)k * kTileElems, sD, kTileElems * 8);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        bulk_s2g(tile_ptr(e.M, T, k, k), sL, kTileElems * 8);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    MEDGP_PHASE(14)
    // forward solve, block k: z_k = X_kk rhs_k (in place)
    tile_matvec_rhs(e, sD, e.rhs + k * MEDGP_NB, e.rhs + k * MEDGP_NB, true, false, gjb->d[0], true, v0);
    MEDGP_PHASE(15)
    if (publish) {
        // A consumer inside the SAME kernel waits for this flag (panel roles of k_potrf_step /
        // k_potrf_flow): it needs X_kk landed and ordered before the generic-proxy flag, and z_k.
        // X_kk^T and L_kk, which nobody needs before the next kernel, go out afterwards.
        if (tid == 0) {
            asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");  // the X_kk store has completed
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        __threadfence();  // z_k
        __syncthreads();
        if (tid == 0) flag_release(publish, publish_value);
    }
    // write back: L_kk and dinv (X, column-major) are whole-tile copies and leave through the TMA
    // engine (one 34816 B bulk store each, issued before the forward solve above); dinvT (X^T)
    // goes out from registers meanwhile, 16 bytes per store.
    double *XTk = e.dinvT + (size_t)k * kTileElems;
    {
        // X^T by 2x2 blocks: a warp reads 4 column pairs x 8 row pairs of X (conflict-free 16-byte
        // loads) and writes 8 column pairs x 64 contiguous bytes of X^T
        const int lane = tid & 31, warp = tid >> 5, al = lane >> 3, bl = lane & 7;
        for (int it = warp; it < 32; it += MEDGP_DIAG_THREADS / 32) {
            const int a = 4 * (it >> 2) + al, b = 8 * (it & 3) + bl;  // X rows 2b, 2b+1; columns 2a, 2a+1
            const double2 q0 = *reinterpret_cast<const double2 *>(sD + (2 * a) * MEDGP_SLD + 2 * b);
            const double2 q1 = *reinterpret_cast<const double2 *>(sD + (2 * a + 1) * MEDGP_SLD + 2 * b);
            *reinterpret_cast<double2 *>(XTk + (2 * b) * MEDGP_SLD + 2 * a) = make_double2(q0.x, q1.x);
            *reinterpret_cast<double2 *>(XTk + (2 * b + 1) * MEDGP_SLD + 2 * a) = make_double2(q0.y, q1.y);
        }
    }
    if (tid == 0) {
        e.blk[k] = gjb->logdet;
        if (*s_fail) fail[e.out_index] = 1;
        // the shared tiles must outlive the bulk stores' reads (at the kernel boundary that is enough)
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (publish) {
        // second signal (value + 1): X_kk^T and L_kk are in place too (the dataflow kernel's inverse roles read X_kk^T)
        if (tid == 0) {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) flag_release(publish, publish_value + 1);
    }
    MEDGP_PHASE(16)
}

// Stand-alone diagonal kernel, one CTA per evaluation: D = K_kk - sum_{l0 <= l < l0+depth} L_kl L_kl^T.
// Left-looking: l0 = 0, depth = k.  Panel-blocked right-looking: l0 = first block column of the
// current panel (everything left of it has already been applied by the trailing updates).
__global__ void __launch_bounds__(MEDGP_DIAG_THREADS, 3)
k_potrf_diag(const EvalDesc *__restrict__ descs, int k, int depth, int *__restrict__ fail, int l0 = 0)
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
                          A = tile_ptr(M, T, k, l0 + l);
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
k_potrf_panel(const EvalDesc *__restrict__ descs, int k, int depth, int fold_diag, int *__restrict__ fail, int l0 = 0)
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
                      A = tile_ptr(M, T, i, l0 + l);
                      B = tile_ptr(M, T, k, l0 + l);
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
k_potrf_step(const EvalDesc *__restrict__ descs, int k, int *__restrict__ fail, int *__restrict__ ticket,
             int l0 = 0, int fold = 1)
{
    // fold = 1 (left-looking, l0 = 0): every panel role folds its tile into its own diagonal block,
    // so the diagonal role has no product to do.  fold = 0 (panel-blocked right-looking, l0 = first
    // block column of the current panel): the diagonal role forms the k - l0 products of its block
    // itself (at most W - 1), the trailing update takes care of everything right of the panel.
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
        const bool products = !fold && k > l0;
        if (products) {
            prefetch_tile_l2(tile_ptr(M, T, k, k));
            gemm_bars_init(&bars);
            double acc[4][4][2];
            acc_zero(acc);
            gemm_nt_tiles(acc, k - l0,
                          [&](int l, const double *&A, const double *&B) {
                              A = tile_ptr(M, T, k, l0 + l);
                              B = A;
                          },
                          smem, &bars, NoStageFn(),
                          [](int, int wm, int wn) { return wm == 0 && wn == 1; },  // lower part only
                          TileEdge{rows_valid(e, k), rows_valid(e, k), MEDGP_NB});
            __syncthreads();  // all warps are done with the ring before it is reused as sP
            acc_to_smem(acc, sP, 1.0);
        }
        __syncthreads();
        diag_block_factor(e, k, sP, sX, &gjb, &s_fail, fail, products, e.flags + k * T + k);
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
    gemm_nt_tiles(acc, k - l0,
                  [&](int l, const double *&A, const double *&B) {
                      A = tile_ptr(M, T, i, l0 + l);
                      B = tile_ptr(M, T, k, l0 + l);
                  },
                  smem, &bars, NoStageFn(), NoSkipFn(), TileEdge{mv, MEDGP_NB, MEDGP_NB});
    acc_rsub_global(acc, Tik);  // P = K_ik - C (does not depend on the diagonal block)
    __syncthreads();            // every warp is done with the pipeline buffers
    acc_to_smem(acc, sP, 1.0);
    if (threadIdx.x == 0)
        while (flag_acquire(e.flags + k * T + k) == 0) __nanosleep(64);
    __syncthreads();
    tile_bulk_g2s(sX, Xk, &bars);
    tile_bulk_wait(&bars);
    __syncthreads();
    gemm2_smem(acc, sP, sX, mv);
    acc_to_global(acc, Tik);
    acc_matvec_rhs(acc, e, e.rhs + k * MEDGP_NB, e.rhs + i * MEDGP_NB, red);
    if (!fold) return;
    // fold into the own diagonal block: K_ii -= L_ik L_ik^T
    acc_to_smem(acc, sP, 1.0);
    __syncthreads();
    syrk_smem(acc, sP, mv);
    double *Kii = tile_ptr(M, T, i, i);
    acc_rsub_global(acc, Kii);
    acc_to_global(acc, Kii);
}

// ------------------------------------------------------------------ dataflow factorisation
// ONE launch for the whole factorisation of every matrix of a sub-chunk (k_potrf_flow) and one for
// the whole triangular inverse (k_trtri_flow).  Every tile is a role; a role streams its tile
// products in order and waits, tile pair by tile pair, for the per-tile flags of its operands, so
// block columns overlap as far as the data dependencies allow: for ONE large matrix the critical
// path shrinks from "every launch of every block column" to diagonal block -> last product ->
// second product per column, while all other products run ahead; for many small matrices the
// launch boundaries (2 per block column) and their tails disappear.
// Roles are tickets drawn from a counter when a CTA starts running, in an order that is a
// topological order of the dependency graph (potrf: by block column, diagonal roles of a column
// first; trtri: by block row of L^-1).  A role only ever waits for roles with smaller tickets, and
// a ticket is only ever held by a CTA that is already running, so forward progress does not depend
// on dispatch order, grid size or what else shares the GPU.
// FlowMap: act[t] = evaluations of the sub-chunk with more than t block rows (the descriptors are
// sorted by size, so these are prefixes), base[k] = first ticket of block column / row k.
#define MEDGP_FLOW_TMAX 64
#ifdef MEDGP_X_TRACE  // timing experiment: globaltimer stamps of the critical roles of evaluation 0 (tools/flow_trace.py)
__device__ unsigned long long g_flow_trace[MEDGP_FLOW_TMAX][16];
__device__ __forceinline__ void flow_stamp(int k, int slot)
{
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_flow_trace[k][slot] = t;
    }
}
#define FLOW_STAMP(cond, k, slot) { if (cond) flow_stamp(k, slot); }
#else
#define FLOW_STAMP(cond, k, slot)
#endif
struct FlowMap {
    int Tmax, total;
    int act[MEDGP_FLOW_TMAX + 1];
    int base[MEDGP_FLOW_TMAX + 1];
};

__device__ __forceinline__ void flag_wait(const int *flag, int at_least = 1)
{
    while (flag_acquire(flag) < at_least) __nanosleep(40);
}

// publish a finished tile: every thread's stores, then the flag (call with all threads)
__device__ __forceinline__ void tile_publish(int *flag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) flag_release(flag);
}

// Roles of k_potrf_flow, in ticket order, for block column c = 0, 1, ...:
//   (c == 0)  DIAG0: factor block 0;
//   (c >= 1)  PRE(c+1): D' = K_{c+1,c+1} - sum_{l<c} L_{c+1,l} L_{c+1,l}^T, everything of the next diagonal
//             block but its last term, written back into the tile (diagonal flag value 1);
//   PANEL(i, c), i > c: L_ic; the role of tile (c+1, c) CONTINUES: it still holds L_{c+1,c} on chip, forms the
//             last term from shared memory, subtracts it from D' and factors block c+1 (diagonal flag value 2).
// Without the chaining the next diagonal block would fetch tile (c+1, c) back through L2 for that last
// product: 8.5 us of the 28 us a block column takes (tools/flow_trace.py, profiles/r02_flow_trace.txt).
// Every role waits for smaller tickets only: PRE(c+1) needs block columns < c, PANEL(i, c) needs block
// columns < c and the factor of block c (chained to a PANEL role of column c-1, or DIAG0), the chained
// factorisation needs PRE(c+1), the first ticket(s) of its own column.
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_potrf_flow(const EvalDesc *__restrict__ descs, const __grid_constant__ FlowMap map, int *__restrict__ fail,
             int *__restrict__ ticket, int with_inverse)
{
    // with_inverse: block column c also carries the roles INV(j, c), j < c, of the triangular inverse
    // (U_jc, as in k_trtri_flow), after its PANEL roles: they need block row c of L (columns < c), X_cc
    // (chained to column c-1), X_jj^T (diagonal flag value 3: stored) and U_jl, l < c (INV roles of
    // earlier columns) -- smaller tickets all.  The factorisation's critical chain leaves most of the
    // SMs idle; the inverse fills them instead of running afterwards.
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    __shared__ __align__(16) GjBufs gjb;
    __shared__ double red[2 * MEDGP_NB];
    __shared__ int s_fail, s_role[3];
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        int c = 0;
        while (c + 1 < map.Tmax && map.base[c + 1] <= t) c++;
        int rem = t - map.base[c], r = -1;  // r = -1: DIAG0 / PRE(c+1); r >= 1: PANEL(c + r, c)
        const int nfirst = c == 0 ? map.act[0] : (c + 1 < map.Tmax ? map.act[c + 1] : 0);
        if (rem >= nfirst) {
            rem -= nfirst;
            r = 1;
            while (c + r < map.Tmax && rem >= map.act[c + r]) rem -= map.act[c + r++];
            if (c + r >= map.Tmax) {  // past the PANEL roles: INV(j, c), r = -2 - j
                r = -2 - rem / map.act[c];
                rem = rem % map.act[c];
            }
        }
        s_role[0] = c; s_role[1] = r; s_role[2] = rem;
        s_fail = 0;
    }
    __syncthreads();
    const int c = s_role[0], r = s_role[1];
    const EvalDesc &e = descs[s_role[2]];
    if (e.skip) return;  // (every role of this evaluation returns: nobody waits for it)
    const int T = e.T;
    double *M = e.M;
    int *flags = e.flags;
    double *sP = smem, *sX = smem + kTileElems;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    int waited = -1;  // (lane 0 of the producer warp) last tile pair whose flags have been seen
    const bool tr = (s_role[2] == 0);
    if (r <= -2) {  // ---- INV(j, i): U_ji = -(sum_{l=j}^{i-1} U_jl L_il^T) X_ii^T, U_jj = X_jj^T
        const int i = c, j = -2 - r, nv = rows_valid(e, i);
        const double *XTj = e.dinvT + (size_t)j * kTileElems;
        gemm_nt_tiles(acc, i - j,
                      [&](int l0, const double *&A, const double *&B) {
                          const int l = j + l0;
                          if (l0 > waited) {
                              if (l0 == 0) flag_wait(flags + j * T + j, 3);  // X_jj^T stored
                              else flag_wait(flags + j * T + l);             // U_jl
                              flag_wait(flags + i * T + l);                  // L_il
                              asm volatile("fence.proxy.async;" ::: "memory");
                              waited = l0;
                          }
                          A = (l0 == 0) ? XTj : tile_ptr(M, T, j, l);
                          B = tile_ptr(M, T, i, l);
                      },
                      smem, &bars, NoStageFn(), [](int ch, int wm, int) { return ch < 2 && wm == 1; },
                      TileEdge{MEDGP_NB, nv, MEDGP_NB});
        __syncthreads();
        if (threadIdx.x == 0) flag_wait(flags + i * T + i, 2);  // X_ii
        __syncthreads();
        tile_bulk_g2s(sX, e.dinv + (size_t)i * kTileElems, &bars);
        acc_to_smem(acc, sP, -1.0);
        tile_bulk_wait(&bars);
        __syncthreads();
        gemm2_smem(acc, sP, sX, MEDGP_NB, nv);
        acc_to_global(acc, tile_ptr(M, T, j, i));
        tile_publish(flags + j * T + i);
        return;
    }
    if (r < 0 && c == 0) {  // ---- DIAG0
        diag_block_factor(e, 0, sP, sX, &gjb, &s_fail, fail, false, flags, 2);
        return;
    }
    if (r < 0) {  // ---- PRE(k), k = c + 1: all but the last term of the diagonal block
        const int k = c + 1;
        double *Kkk = tile_ptr(M, T, k, k);
        prefetch_tile_l2(Kkk);
        gemm_nt_tiles(acc, k - 1,
                      [&](int l, const double *&A, const double *&B) {
                          if (l > waited) {
                              flag_wait(flags + k * T + l);
                              asm volatile("fence.proxy.async;" ::: "memory");  // the tile's generic-proxy writes -> our bulk copies
                              waited = l;
                          }
                          A = tile_ptr(M, T, k, l);
                          B = A;
                      },
                      smem, &bars, NoStageFn(), [](int, int wm, int wn) { return wm == 0 && wn == 1; },
                      TileEdge{rows_valid(e, k), rows_valid(e, k), MEDGP_NB});
        acc_rsub_global(acc, Kkk);  // (the skipped upper-right quadrant keeps K: it is masked when the block is factored)
        acc_to_global(acc, Kkk);
        tile_publish(flags + k * T + k);  // value 1: D' is in place
        return;
    }
    // ---- PANEL(i, k): L_ik = (K_ik - sum_{l<k} L_il L_kl^T) X_kk^T
    const int k = c, i = c + r;
    FLOW_STAMP(tr && r == 1, k, 8)  // role started
    double *Tik = tile_ptr(M, T, i, k);
    const double *Xk = e.dinv + (size_t)k * kTileElems;
    prefetch_tile_l2(Tik);
    const int mv = rows_valid(e, i);
    gemm_nt_tiles(acc, k,
                  [&](int l, const double *&A, const double *&B) {
                      if (l > waited) {
                          flag_wait(flags + i * T + l);
                          flag_wait(flags + k * T + l);
                          asm volatile("fence.proxy.async;" ::: "memory");
                          waited = l;
                      }
                      A = tile_ptr(M, T, i, l);
                      B = tile_ptr(M, T, k, l);
                  },
                  smem, &bars, NoStageFn(), NoSkipFn(), TileEdge{mv, MEDGP_NB, MEDGP_NB});
    acc_rsub_global(acc, Tik);  // P = K_ik - C (K_ik comes from the assembly kernel: complete at launch)
    __syncthreads();            // every warp is done with the pipeline buffers
    acc_to_smem(acc, sP, 1.0);
    FLOW_STAMP(tr && r == 1, k, 9)   // products done, waiting for X_kk
    if (threadIdx.x == 0) flag_wait(flags + k * T + k, 2);
    FLOW_STAMP(tr && r == 1, k, 10)  // flag seen
    __syncthreads();
    tile_bulk_g2s(sX, Xk, &bars);
    tile_bulk_wait(&bars);
    __syncthreads();
    FLOW_STAMP(tr && r == 1, k, 11)  // X_kk in shared memory
    gemm2_smem(acc, sP, sX, mv);
    acc_to_global(acc, Tik);
    FLOW_STAMP(tr && r == 1, k, 12)  // second product done, tile stored
    // forward solve: rhs_i -= L_ik z_k.  The roles (i, k') of one block row run in the order of k'
    // (role (i, k) has waited for tile (i, k-1), published after ITS update), so the updates of rhs_i
    // are applied in a fixed order: bit-reproducible, no atomics
    acc_matvec_rhs(acc, e, e.rhs + k * MEDGP_NB, e.rhs + i * MEDGP_NB, red);
    FLOW_STAMP(tr && r == 1, k, 13)  // right-hand sides updated
    tile_publish(flags + i * T + k);
    FLOW_STAMP(tr && r == 1, k, 14)  // published
    if (r != 1) return;
    // ---- chained: block i = k + 1.  Last term L_ik L_ik^T from the tile still on chip, D' from PRE(i)
    acc_to_smem(acc, sP, 1.0);  // (tile_publish ended with a block barrier: the second product is done with sP)
    __syncthreads();
    syrk_smem(acc, sP, mv);
    __syncthreads();  // everyone is done reading sP
    acc_to_smem(acc, sP, 1.0);
    if (k >= 1 && threadIdx.x == 0) flag_wait(flags + i * T + i, 1);  // (block 1 has no earlier term: its tile is K_11)
    __syncthreads();
    FLOW_STAMP(tr, i, 2)  // last term formed, D' available
    diag_block_factor(e, i, sP, sX, &gjb, &s_fail, fail, true, flags + i * T + i, 2);
    FLOW_STAMP(tr, i, 3)  // factor done, flag published, remaining stores issued
}

// the triangular inverse the same way: role (j, i), j < i: U_ji = -(sum_{l=j}^{i-1} U_jl L_il^T) X_ii^T with
// U_jj = X_jj^T; it waits for U_jl, j < l < i (tiles of earlier block rows: smaller tickets).
// base[i] = first ticket of block row i (i >= 1), act[i] evaluations have it, i roles each.
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_trtri_flow(const EvalDesc *__restrict__ descs, const __grid_constant__ FlowMap map, int *__restrict__ ticket)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    __shared__ int s_role[3];
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        int i = 1;
        while (i + 1 < map.Tmax && map.base[i + 1] <= t) i++;
        const int rem = t - map.base[i];
        s_role[0] = i; s_role[1] = rem / map.act[i]; s_role[2] = rem % map.act[i];
    }
    __syncthreads();
    const int i = s_role[0], j = s_role[1];
    const EvalDesc &e = descs[s_role[2]];
    if (e.skip) return;
    const int T = e.T, nv = rows_valid(e, i);
    double *M = e.M;
    int *flags = e.flags;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    const double *XTj = e.dinvT + (size_t)j * kTileElems;
    int waited = 0;  // (lane 0 of the producer warp) the first operand is X_jj^T, final since the factorisation
    gemm_nt_tiles(acc, i - j,
                  [&](int l0, const double *&A, const double *&B) {
                      const int l = j + l0;
                      if (l0 > waited) {
                          flag_wait(flags + j * T + l);
                          asm volatile("fence.proxy.async;" ::: "memory");
                          waited = l0;
                      }
                      A = (l0 == 0) ? XTj : tile_ptr(M, T, j, l);
                      B = tile_ptr(M, T, i, l);
                  },
                  smem, &bars, NoStageFn(), [](int ch, int wm, int) { return ch < 2 && wm == 1; },
                  TileEdge{MEDGP_NB, nv, MEDGP_NB});
    __syncthreads();
    double *sP = smem, *sX = smem + kTileElems;
    tile_bulk_g2s(sX, e.dinv + (size_t)i * kTileElems, &bars);
    acc_to_smem(acc, sP, -1.0);
    tile_bulk_wait(&bars);
    __syncthreads();
    gemm2_smem(acc, sP, sX, MEDGP_NB, nv);
    acc_to_global(acc, tile_ptr(M, T, j, i));
    tile_publish(flags + j * T + i);
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
// matrix, each k tiles deep).  The right-looking form exposes (T-k)^2/2 independent products
// per step instead, at the price of re-reading the trailing tiles once per update.  With one
// update per block column those products are one tile deep and the update is bound by HBM / L2
// traffic (4 tiles moved per tile product) as soon as the matrices in flight outgrow the L2, so
// the factorisation is PANEL-BLOCKED: W block columns are factored left-looking among
// themselves (k_potrf_diag / k_potrf_panel with l0 = first column of the panel), then applied
// to the trailing matrix in one W-tile-deep update:
//   potrf:  K_ij -= sum_{l in panel} L_il L_jl^T   for panel end <= j <= i   (k_syrk_update)
//   trtri:  Acc_ji (+)= U_jk L_ik^T        for j <= k < i      (k_trtri_update), then
//           U_j,k+1 = -Acc_j,k+1 X_k+1^T                       (k_trtri_row, right_looking = 1)
__global__ void __launch_bounds__(MEDGP_GEMM_THREADS, 3)
k_syrk_update(const EvalDesc *__restrict__ descs, int k0, int nk, int j0, int ncol)
{
    // Trailing update with the nk block columns k0 .. k0+nk-1 of L (a finished panel):
    //   K_ij -= sum_{l} L_il L_jl^T   for the lower tiles (i, j) of block columns j0 .. j0+ncol-1.
    // ncol == 0: every block column from j0 on (lower-triangle enumeration); otherwise blockIdx.x
    // = (row offset) * ncol + (column offset), row offset counted from the tile's own diagonal.
    // Look-ahead schedule: the block columns of the NEXT panel are updated on the evaluation's
    // main stream (the critical path), everything right of them on an auxiliary stream beside
    // the next panel's factorisation.
    extern __shared__ __align__(128) double smem[];
    __shared__ GemmBars bars;
    const EvalDesc &e = descs[blockIdx.y];
    int a, b;
    if (ncol == 0) {
        tri_index(blockIdx.x, a, b);
    } else {
        b = blockIdx.x % ncol;
        a = b + blockIdx.x / ncol;
    }
    const int i = j0 + a, j = j0 + b, T = e.T;
    if (i >= T || e.skip) return;
    gemm_bars_init(&bars);
    double acc[4][4][2];
    acc_zero(acc);
    double *M = e.M;
    const bool diag = (i == j);
    gemm_nt_tiles(acc, nk,
                  [&](int l, const double *&A, const double *&B) {
                      A = tile_ptr(M, T, i, k0 + l);
                      B = tile_ptr(M, T, j, k0 + l);
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
