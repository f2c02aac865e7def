// common.cuh -- shared definitions for libmedgp_cuda.so (sm_100a only).
//
// HBM layout of one in-flight evaluation (all FP64):
//   M      T x T tiles (T = npad/64, npad = n rounded up to 64, padded with the identity), stored
//          TILE-MAJOR: tile (ti, tj) is the contiguous block number tj*T + ti of 64 columns x
//          pitch 68 doubles (34816 B).  The pitch is the bank-conflict-free shared-memory pitch
//          of the DMMA fragment loads, so a 16-column k-panel of a tile is ONE contiguous
//          8704 B bulk copy straight into its pipeline stage.  Lower tiles hold K -> L -> K^-1
//          in place; strictly-upper tiles hold U = (L^-1)^T once the triangular inverse ran.
//   dinv   T tiles (same 64 x pitch-68 format): X_kk = inv(L_kk) (lower, zeros above)
//   dinvT  the same blocks transposed (X_kk^T, upper)
//   rhs    nrhs x npad: row 0 = y -> z = L^-1 y ; rows 1.. = k* -> L^-1 k* (prediction)
//   alpha  npad: K^-1 y
//   cs     Q x npad x (cos, sin)(2 PI mu_q t_i)
//   par    derived hyper-parameters (ParLayout below)
//   part   gradient partial sums, one row of (3Q+1) doubles per work item
//   blk    T per-block partial log-determinants
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MEDGP_NB 64        // tile edge (rows/cols of a tile, depth of one tile product)
#define MEDGP_SLD 68       // shared-memory pitch of a 64-row tile column (== 4 mod 16: DMMA fragment loads are bank-conflict free)
#define MEDGP_KC 16        // k-depth of one pipeline stage
#define MEDGP_NSTAGE 4     // pipeline stages
#define MEDGP_QMAX 8       // compile-time bound on mixture components
#define MEDGP_GEMM_THREADS 128

struct ModelDims {
    int Q, D, R, P;
    double pi;
    // offsets (in doubles) inside the per-evaluation parameter block
    int oB, oSig2, oW, oC, oA, oKappa, oBdiag, parLen;
};

struct __align__(16) EvalDesc {
    double *M, *dinv, *dinvT, *rhs, *alpha, *cs, *par, *part, *blk;
    int *flags;              // T ints: flags[k] = 1 once diagonal block k is factored (k_potrf_step)
    const double *t, *y;
    const int *meta, *off;
    const int4 *items;
    const int *seg_start;
    const double *star_t;    // prediction points of this evaluation (device), or null
    const int *star_meta;
    int n, npad, T, nitems;
    int jitter, nrhs, nstar, out_index;
    int star_out;            // offset of this evaluation's predictions in the output arrays
    int skip;                // device-side jitter loop (k_retry_decide): 1 = this evaluation is final, every kernel passes over it
    double trange2;          // (max t - min t)^2 of the series: bounds every tau^2 of the training block
    const int *gstart;       // time-ordered series: first point of every same-timestamp group (ngroups + 1)
    const int *perm;         // internal position -> caller's point index
    int ngroups, pad1;
    const int2 *frange;      // per 64-point block of the series: (lowest, highest) feature among its points
};

// tile-major addressing (kTileElems doubles per tile, column pitch MEDGP_SLD)
__host__ __device__ __forceinline__ size_t tile_off(int T, int ti, int tj)
{
    return ((size_t)tj * T + ti) * (size_t)(MEDGP_NB * MEDGP_SLD);
}
__host__ __device__ __forceinline__ size_t elem_off(int T, int i, int j)
{
    return tile_off(T, i >> 6, j >> 6) + (size_t)((j & 63) * MEDGP_SLD + (i & 63));
}

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA engine, UBLKCP) + FP64 tensor-core MMA (DMMA)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar`.
// src, dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 1-D bulk copy shared -> global through the TMA engine (bulk async-group completion).  The
// caller orders its shared-memory writes with fence.proxy.async + a barrier first, commits the
// group and waits for it (at least for its reads) before the shared source is reused or freed.
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}

// inter-CTA flag (same kernel): release by the producer CTA, acquire-spin by the consumers
__device__ __forceinline__ void flag_release(int *flag)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}
__device__ __forceinline__ int flag_acquire(const int *flag)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}

// Warp -> quadrant assignment of the tile GEMM kernels.  Wherever a quadrant of a tile product is
// skipped (zero blocks of triangular operands, the unused part of symmetric results, rows beyond
// the end of a ragged matrix) the same warp indices idle in every CTA.  If warp w of every CTA sat
// on SM sub-partition w, rotating the assignment by the block index would spread the skipped
// quadrants over the four FP64 pipes; measured on B200 (MEDGP_WARP_ROT=1 against 0 on the C3
// cohort): no difference, so the plain assignment is the default and the switch stays for experiments.
#ifndef MEDGP_WARP_ROT
#define MEDGP_WARP_ROT 0
#endif
__device__ __forceinline__ int gemm_warp()
{
#if MEDGP_WARP_ROT
    return ((threadIdx.x >> 5) + blockIdx.x + blockIdx.y) & 3;
#else
    return threadIdx.x >> 5;
#endif
}

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// per-thread asynchronous global->shared copies (LDGSTS) for warp-private double buffering
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem)
{
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem)), "l"(gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------
// exp(x) for x <= 0, NV values at once.  The covariance kernels need one exp per mixture
// component per point pair; libdevice's exp() gets serialised per component, which leaves the
// FP64 pipe waiting on its own dependent chains.  Writing the NV range reductions and Horner
// steps side by side gives the scheduler NV independent chains.
// x = k (ln2/64) + r, |r| <= ln2/128 (Cody-Waite, k*HI exact for |k| < 2^20), so
// exp(x) = 2^(k>>6) * 2^((k&63)/64) * exp(r): a 64-entry table (correctly rounded, staged in
// shared memory) and a degree-5 Taylor polynomial (truncation 3.5e-17 relative, below half an
// ulp); the power of two goes straight into the exponent field.  x < -708 flushes to 0 (true
// value below 1e-307).  10 FP64 operations per value; max observed error vs expl(): 1.5 ulp.
// ---------------------------------------------------------------------------------------
#define MEDGP_EXP_TAB 64
__constant__ double c_exp2_tab[MEDGP_EXP_TAB] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

__device__ __forceinline__ void exp_tab_stage(double *s_tab)  // call by the whole CTA, then barrier
{
    for (int i = threadIdx.x; i < MEDGP_EXP_TAB; i += blockDim.x) s_tab[i] = c_exp2_tab[i];
}

// polynomial and reduction constants live in constant memory so that the DFMAs take them as
// constant-bank operands instead of re-materialising 64-bit immediates in registers
__constant__ double c_expc[9] = {
    92.33248261689366,        // 0: 64/ln2
    0.010830424695086549,     // 1: ln2/64 high part (low 20 mantissa bits zero)
    1.162596423439437e-12,    // 2: ln2/64 low part
    8.33333333333333333333e-03, 4.16666666666666666667e-02, 1.66666666666666666667e-01,
    0.5, 1.0, 1.0};           // 3..8: 1/5! .. 1/0!

// Largest |x| for which the unchecked variant is valid (k = x*32/ln2 must fit 32 bits).
#define MEDGP_EXP_UNCHECKED_MAX 2.0e7

// CHECKED: any x <= 0 (x < -708 gives exactly 0).  !CHECKED: requires x >= -MEDGP_EXP_UNCHECKED_MAX;
// results below 2^-1022 come out as some value < 2^-1021 instead of exactly 0 (the exponent
// field saturates at 0), which saves a compare and two selects per value.
template <int NV, bool CHECKED = true>
__device__ __forceinline__ void exp_nonpos(const double (&x)[NV], double (&out)[NV], const double *s_tab)
{
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    double r[NV], p[NV];
    int k[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const double t = fma(x[i], c_expc[0], MAGIC);
        k[i] = __double2loint(t);
        const double kf = t - MAGIC;
        r[i] = fma(-kf, c_expc[1], x[i]);
        r[i] = fma(-kf, c_expc[2], r[i]);
        p[i] = c_expc[3];
    }
#pragma unroll
    for (int c = 4; c < 9; c++)
#pragma unroll
        for (int i = 0; i < NV; i++) p[i] = fma(p[i], r[i], c_expc[c]);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const double tj = s_tab[k[i] & (MEDGP_EXP_TAB - 1)];
        int m = k[i] >> 6;
        if (!CHECKED) m = max(m, -1023);  // exponent field saturates at 0
        const double v = p[i] * __hiloint2double(__double2hiint(tj) + (m << 20), __double2loint(tj));
        out[i] = (CHECKED && x[i] < -708.0) ? 0.0 : v;
    }
}

// ---------------------------------------------------------------------------------------
// Tile GEMM core:  C(64x64) = sum_l A_l * B_l^T,  A_l, B_l 64x64 column-major tiles in HBM.
// One CTA of 4 warps; warp (wm, wn) owns the 32x32 quadrant, as 4x4 DMMA 8x8 sub-tiles:
//   acc[a][b][e] = C[32wm + 8a + lane/4][32wn + 8b + 2(lane%4) + e]
// Operands are staged one 16-column panel (8704 B, contiguous in the tile-major HBM layout) per
// cp.async.bulk into a 4-stage ring of [k][m] / [k][n] panels with pitch MEDGP_SLD, guarded by
// full/empty mbarriers.
// ---------------------------------------------------------------------------------------
constexpr int kPanelElems = MEDGP_KC * MEDGP_SLD;            // one operand panel of a stage
constexpr int kStageElems = 2 * kPanelElems;                 // A panel + B panel
constexpr int kTileElems = MEDGP_NB * MEDGP_SLD;             // a full 64x64 tile with pitch
constexpr int kGemmSmemBytes = MEDGP_NSTAGE * kStageElems * 8; // 69632 B == 2 full tiles

struct GemmBars {
    uint64_t full[MEDGP_NSTAGE];
    uint64_t empty[MEDGP_NSTAGE];
    uint64_t aux;  // whole-tile bulk copy of an epilogue operand (X_kk)
};

__device__ __forceinline__ void gemm_bars_init(GemmBars *bars)
{
    if (threadIdx.x == 0) {
        for (int s = 0; s < MEDGP_NSTAGE; s++) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], MEDGP_GEMM_THREADS / 32);
        }
        mbar_init(&bars->aux, 1);
        mbar_fence_init();
    }
    __syncthreads();
}

__device__ __forceinline__ void acc_zero(double (&acc)[4][4][2])
{
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
}

// accumulate `depth4` k-steps of 4 from panels sA ([k][m], pitch SLD) and sB ([k][n])
__device__ __forceinline__ void mma_panels(double (&acc)[4][4][2], const double *sA,
                                           const double *sB, int depth4, int wm, int wn, int lane)
{
#ifdef MEDGP_X_NODMMA  // timing experiment: the floor set by everything but the DMMAs (results are wrong)
    return;
#endif
    const int r = lane >> 2, kq = lane & 3;
    const double *pa = sA + kq * MEDGP_SLD + wm * 32 + r;
    const double *pb = sB + kq * MEDGP_SLD + wn * 32 + r;
#pragma unroll 4
    for (int kk = 0; kk < depth4; kk++) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            a[u] = pa[8 * u];
            b[u] = pb[8 * u];
        }
#pragma unroll
        for (int x = 0; x < 4; x++)
#pragma unroll
            for (int y = 0; y < 4; y++) dmma884(acc[x][y][0], acc[x][y][1], a[x], b[y]);
        pa += 4 * MEDGP_SLD;
        pb += 4 * MEDGP_SLD;
    }
}

// the same for a tile at the ragged end of a matrix: only the first mx (ny) 8-row sub-tiles of this
// warp's quadrant hold rows (columns) below n; the others are known to be zero and stay untouched
__device__ __forceinline__ void mma_panels_edge(double (&acc)[4][4][2], const double *sA,
                                                const double *sB, int depth4, int wm, int wn, int lane,
                                                int mx, int ny)
{
    const int r = lane >> 2, kq = lane & 3;
    const double *pa = sA + kq * MEDGP_SLD + wm * 32 + r;
    const double *pb = sB + kq * MEDGP_SLD + wn * 32 + r;
#pragma unroll 2
    for (int kk = 0; kk < depth4; kk++) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            a[u] = pa[8 * u];
            b[u] = pb[8 * u];
        }
#pragma unroll
        for (int x = 0; x < 4; x++)
            if (x < mx) {
#pragma unroll
                for (int y = 0; y < 4; y++)
                    if (y < ny) dmma884(acc[x][y][0], acc[x][y][1], a[x], b[y]);
            }
        pa += 4 * MEDGP_SLD;
        pb += 4 * MEDGP_SLD;
    }
}

// Ragged end of a matrix (n is not a multiple of 64; the padding is the identity, so the
// off-diagonal tiles of the last block row / column are zero beyond n): m = rows of the A tiles
// that can be non-zero, n = rows of the B tiles, klast = columns of the LAST tile pair of the sum.
struct TileEdge {
    int m = MEDGP_NB, n = MEDGP_NB, klast = MEDGP_NB;
};
// number of 8-row sub-tiles of quadrant w (rows 32w ..) that start below `valid`
__device__ __forceinline__ int edge_subtiles(int valid, int w)
{
    return min(4, max(0, (valid - 32 * w + 7) >> 3));
}

struct NoStageFn {
    __device__ __forceinline__ void operator()(int, const double *) const {}
};
struct NoSkipFn {
    __device__ __forceinline__ bool operator()(int, int, int) const { return false; }
};

// TileFn: void operator()(int l, const double*& A, const double*& B) -> tile base pointers.
// StageFn: void operator()(int ch, const double *stage): extra work of every thread on the
// resident k-panel pair ch (A panel at stage, B panel at stage + kPanelElems, [k][row] with pitch
// SLD) before it is handed back to the copy engine.
// SkipFn: bool operator()(int ch, int wm, int wn): true when the 32x32 quadrant (wm, wn) of this
// warp gets nothing from k-panel ch (a zero block of a triangular operand): its DMMAs are skipped.
template <int NST = MEDGP_NSTAGE, class TileFn, class StageFn = NoStageFn, class SkipFn = NoSkipFn>
__device__ __forceinline__ void gemm_nt_tiles(double (&acc)[4][4][2], int nl, TileFn tiles,
                                              double *smem, GemmBars *bars, StageFn stagefn = StageFn(),
                                              SkipFn skipfn = SkipFn(), TileEdge edge = TileEdge())
{
    const int warp = gemm_warp(), lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = warp >> 1;
    const int nch = nl * (MEDGP_NB / MEDGP_KC);
    const int mx = edge_subtiles(edge.m, wm), ny = edge_subtiles(edge.n, wn);
    const bool ragged = (mx < 4) || (ny < 4);
    constexpr uint32_t kPanelBytes = kPanelElems * 8;  // 8704 B, one bulk copy

    auto issue = [&](int ch) {
        if (lane == 0) {
            const int s = ch % NST;
            const int l = ch / (MEDGP_NB / MEDGP_KC), cc = ch % (MEDGP_NB / MEDGP_KC);
            const double *A, *B;
            tiles(l, A, B);
#ifdef MEDGP_X_NOLOAD  // timing experiment: the floor set by the DMMAs alone (operands are never fetched)
            mbar_arrive(&bars->full[s]);
#else
            mbar_arrive_expect_tx(&bars->full[s], 2 * kPanelBytes);
            double *stage = smem + s * kStageElems;
            bulk_g2s(stage, A + cc * kPanelElems, kPanelBytes, &bars->full[s]);
            bulk_g2s(stage + kPanelElems, B + cc * kPanelElems, kPanelBytes, &bars->full[s]);
#endif
        }
        __syncwarp();
    };

    if (warp == 0)
        for (int ch = 0; ch < NST - 1 && ch < nch; ch++) issue(ch);

    for (int ch = 0; ch < nch; ch++) {
        const int s = ch % NST;
        if (warp == 0) {
            const int nxt = ch + NST - 1;
            if (nxt < nch) {
                if (nxt >= NST)
                    mbar_wait(&bars->empty[nxt % NST], ((nxt / NST) - 1) & 1);
                issue(nxt);
            }
        }
        mbar_wait(&bars->full[s], (ch / NST) & 1);
        const double *stage = smem + s * kStageElems;
        int d4 = MEDGP_KC / 4;
        if (edge.klast < MEDGP_NB && ch >= nch - MEDGP_NB / MEDGP_KC)  // k-steps of the last tile pair below klast
            d4 = min(d4, max(0, (edge.klast - MEDGP_KC * (ch & (MEDGP_NB / MEDGP_KC - 1)) + 3) >> 2));
        if (!skipfn(ch, wm, wn)) {
            if (ragged || d4 < MEDGP_KC / 4)
                mma_panels_edge(acc, stage, stage + kPanelElems, d4, wm, wn, lane, mx, ny);
            else
                mma_panels(acc, stage, stage + kPanelElems, MEDGP_KC / 4, wm, wn, lane);
        }
        stagefn(ch, stage);
        // The stage is about to be handed back to the async proxy (bulk copy): order this
        // thread's generic-proxy reads before it.  Without the fence the refill was observed
        // to race with the fragment loads under load (1e-3 errors, run-to-run differences).
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->empty[s]);
    }
}

// write the accumulator tile (times sign) into a pitch-SLD column-major shared tile
__device__ __forceinline__ void acc_to_smem(const double (&acc)[4][4][2], double *sT, double sign)
{
    const int warp = gemm_warp(), lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = warp >> 1, r = lane >> 2, kq = lane & 3;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int row = wm * 32 + 8 * a + r, col = wn * 32 + 8 * b + 2 * kq;
            sT[col * MEDGP_SLD + row] = sign * acc[a][b][0];
            sT[(col + 1) * MEDGP_SLD + row] = sign * acc[a][b][1];
        }
}

// acc = base - acc, base a pitch-SLD HBM tile
__device__ __forceinline__ void acc_rsub_global(double (&acc)[4][4][2], const double *G)
{
    const int warp = gemm_warp(), lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = warp >> 1, r = lane >> 2, kq = lane & 3;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int row = wm * 32 + 8 * a + r, col = wn * 32 + 8 * b + 2 * kq;
            acc[a][b][0] = G[col * MEDGP_SLD + row] - acc[a][b][0];
            acc[a][b][1] = G[(col + 1) * MEDGP_SLD + row] - acc[a][b][1];
        }
}

__device__ __forceinline__ void acc_to_global(const double (&acc)[4][4][2], double *G)
{
    const int warp = gemm_warp(), lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = warp >> 1, r = lane >> 2, kq = lane & 3;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int row = wm * 32 + 8 * a + r, col = wn * 32 + 8 * b + 2 * kq;
            G[col * MEDGP_SLD + row] = acc[a][b][0];
            G[(col + 1) * MEDGP_SLD + row] = acc[a][b][1];
        }
}

// copy a pitch-SLD HBM tile into a pitch-SLD shared tile (contiguous 34816 B)
__device__ __forceinline__ void tile_g2s_plain(double *sT, const double *G)
{
    for (int idx = threadIdx.x; idx < kTileElems / 2; idx += blockDim.x)
        reinterpret_cast<double2 *>(sT)[idx] = reinterpret_cast<const double2 *>(G)[idx];
}

// the same through the TMA engine: one 34816 B bulk copy, completion on bars->aux (phase 0).
// Call after a block barrier that retired every earlier use of sT; wait with tile_bulk_wait.
__device__ __forceinline__ void tile_bulk_g2s(double *sT, const double *G, GemmBars *bars)
{
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");  // generic -> async, shared and global
        mbar_arrive_expect_tx(&bars->aux, kTileElems * 8);
        bulk_g2s(sT, G, kTileElems * 8, &bars->aux);
    }
}
__device__ __forceinline__ void tile_bulk_wait(GemmBars *bars) { mbar_wait(&bars->aux, 0); }

// pull a tile towards L2 ahead of its use (epilogue operands of the panel kernels)
__device__ __forceinline__ void prefetch_tile_l2(const double *G)
{
    const char *p = reinterpret_cast<const char *>(G);
    for (int off = threadIdx.x * 128; off < kTileElems * 8; off += blockDim.x * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

// block-wide sum of `NV` doubles per thread; result valid in thread 0.  scratch: NV*32 doubles.
template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        v[i] = x;
    }
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; i++) scratch[i * 32 + warp] = v[i];
    __syncthreads();
    if (threadIdx.x == 0)
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double s = 0.0;
            for (int w = 0; w < nwarp; w++) s += scratch[i * 32 + w];
            v[i] = s;
        }
    __syncthreads();
}
