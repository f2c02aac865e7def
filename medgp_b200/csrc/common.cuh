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
    int *flags;              // T x T ints: flags[i*T + j] = 1 once tile (i, j) is final (lower: L, upper: U = L^-T).
                             // k_potrf_step uses the diagonal entries only, the dataflow kernels all of them
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
__device__ __forceinline__ void flag_release(int *flag, int value = 1)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
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
// x = k (ln2/256) + r, |r| <= ln2/512 (Cody-Waite, k*HI exact for |k| < 2^20, i.e. |x| < 2839), so
// exp(x) = 2^(k>>8) * 2^((k&255)/256) * exp(r): a 256-entry table (correctly rounded, staged in
// shared memory) and a degree-4 Taylor polynomial (truncation r^5/120 <= 3.8e-17 relative, below
// half an ulp); the power of two goes straight into the exponent field.  x < -708 flushes to 0
// (true value below 1e-307).  9 FP64 operations per value; error against expl() over 20 M samples of
// [-708, 0]: 2.3 ulp maximum, 0.40 ulp mean (tools/exp_ulp.cu).
// ---------------------------------------------------------------------------------------
#define MEDGP_EXP_TAB 256
// 2^(i/256), correctly rounded (generated with 60-digit decimal arithmetic).  A __device__ array, not
// __constant__: the CTA stages it with one coalesced load per thread (divergent constant reads serialise).
__device__ const double g_exp2_tab[MEDGP_EXP_TAB] = {
    1.0, 1.0027112750502025, 1.0054299011128027, 1.0081558981184175,
    1.0108892860517005, 1.0136300849514894, 1.016378314910953, 1.019133996077738,
    1.0218971486541166, 1.0246677928971357, 1.0274459491187637, 1.030231637686041,
    1.0330248790212284, 1.0358256936019572, 1.0386341019613787, 1.041450124688316,
    1.0442737824274138, 1.0471050958792898, 1.0499440858006872, 1.0527907730046264,
    1.0556451783605572, 1.0585073227945128, 1.061377227289262, 1.0642549128844645,
    1.0671404006768237, 1.0700337118202419, 1.0729348675259756, 1.075843889062791,
    1.0787607977571199, 1.0816856149932152, 1.0846183622133092, 1.0875590609177697,
    1.0905077326652577, 1.0934643990728858, 1.0964290818163769, 1.099401802630222,
    1.102382583307841, 1.1053714457017412, 1.1083684117236787, 1.1113735033448175,
    1.1143867425958924, 1.1174081515673693, 1.1204377524096067, 1.12347556733302,
    1.1265216186082418, 1.129575928566288, 1.1326385195987192, 1.1357094141578055,
    1.1387886347566916, 1.1418762039695616, 1.1449721444318042, 1.148076478840179,
    1.1511892299529827, 1.154310420590216, 1.1574400736337511, 1.1605782120274988,
    1.1637248587775775, 1.1668800369524817, 1.1700437696832502, 1.1732160801636373,
    1.1763969916502812, 1.1795865274628758, 1.182784710984341, 1.1859915656609938,
    1.189207115002721, 1.1924313825831512, 1.1956643920398273, 1.1989061670743806,
    1.202156731452703, 1.2054161090051239, 1.2086843236265816, 1.2119613992768012,
    1.215247359980469, 1.2185422298274085, 1.2218460329727576, 1.2251587936371455,
    1.22848053610687, 1.2318112847340759, 1.2351510639369334, 1.2384998981998165,
    1.241857812073484, 1.245224830175258, 1.2486009771892048, 1.2519862778663162,
    1.255380757024691, 1.2587844395497165, 1.2621973503942507, 1.2656195145788063,
    1.2690509571917332, 1.2724917033894028, 1.275941778396392, 1.2794012075056693,
    1.2828700160787783, 1.2863482295460256, 1.2898358734066657, 1.2933329732290895,
    1.2968395546510096, 1.3003556433796506, 1.3038812651919358, 1.3074164459346773,
    1.3109612115247644, 1.3145155879493546, 1.318079601266064, 1.3216532776031575,
    1.3252366431597413, 1.3288297242059544, 1.3324325470831615, 1.3360451382041458,
    1.339667524053303, 1.3432997311868353, 1.3469417862329458, 1.3505937158920345,
    1.3542555469368927, 1.3579273062129011, 1.3616090206382248, 1.365300717204012,
    1.3690024229745905, 1.3727141650876684, 1.3764359707545302, 1.380167867260238,
    1.383909881963832, 1.387662042298529, 1.3914243757719262, 1.3951969099662003,
    1.3989796725383112, 1.4027726912202048, 1.4065759938190154, 1.4103896082172707,
    1.4142135623730951, 1.4180478843204152, 1.4218926021691656, 1.4257477441054942,
    1.42961333839197, 1.433489413367789, 1.4373759974489824, 1.4412731191286257,
    1.4451808069770467, 1.449099089642035, 1.4530279958490526, 1.4569675544014438,
    1.460917794180647, 1.4648787441464057, 1.4688504333369818, 1.4728328908693675,
    1.4768261459394993, 1.4808302278224719, 1.4848451658727524, 1.488870989524397,
    1.4929077282912648, 1.4969554117672355, 1.5010140696264256, 1.5050837316234065,
    1.5091644275934228, 1.5132561874526098, 1.5173590411982147, 1.5214730189088146,
    1.5255981507445384, 1.529734466947287, 1.533881997840956, 1.5380407738316568,
    1.5422108254079407, 1.5463921831410214, 1.550584877685, 1.5547889397770887,
    1.559004400237837, 1.5632312899713576, 1.567469639965553, 1.5717194812923414,
    1.5759808451078865, 1.5802537626528246, 1.5845382652524937, 1.588834384317164,
    1.593142151342267, 1.597461597908627, 1.6017927556826934, 1.606135656416771,
    1.6104903319492543, 1.6148568142048607, 1.6192351351948637, 1.6236253270173289,
    1.6280274218573478, 1.632441451987275, 1.6368674497669644, 1.6413054476440063,
    1.645755478153965, 1.6502175739206177, 1.6546917676561943, 1.6591780921616162,
    1.6636765803267364, 1.6681872651305825, 1.6727101796415966, 1.6772453570178785,
    1.681792830507429, 1.6863526334483934, 1.6909247992693053, 1.6955093614893326,
    1.7001063537185235, 1.7047158096580513, 1.709337763100463, 1.713972247929926,
    1.718619298122478, 1.723278947746274, 1.7279512309618377, 1.732636182022311,
    1.7373338352737062, 1.7420442251551564, 1.746767386199169, 1.7515033530318782,
    1.7562521603732995, 1.761013843037584, 1.7657884359332727, 1.7705759740635547,
    1.7753764925265212, 1.7801900265154245, 1.785016611318935, 1.789856282321401,
    1.7947090750031072, 1.7995750249405351, 1.804454167806624, 1.809346539371032,
    1.8142521755003989, 1.8191711121586085, 1.8241033854070534, 1.8290490314048973,
    1.8340080864093424, 1.8389805867758937, 1.843966568958626, 1.8489660695104508,
    1.8539791250833855, 1.8590057724288205, 1.864046048397789, 1.8690999899412386,
    1.8741676341103, 1.8792490180565602, 1.8843441790323345, 1.8894531543909392,
    1.8945759815869656, 1.8997126981765553, 1.9048633418176741, 1.9100279502703899,
    1.9152065613971474, 1.9203992131630474, 1.925605943636125, 1.930826790987627,
    1.9360617934922943, 1.9413109895286405, 1.9465744175792332, 1.9518521162309783,
    1.9571441241754002, 1.9624504802089273, 1.9677712232331759, 1.9731063922552343,
    1.978456026387951, 1.9838201648502194, 1.9891988469672663, 1.9945921121709402};

__device__ __forceinline__ void exp_tab_stage(double *s_tab)  // call by the whole CTA, then barrier
{
    for (int i = threadIdx.x; i < MEDGP_EXP_TAB; i += blockDim.x) s_tab[i] = __ldg(g_exp2_tab + i);
}

// polynomial and reduction constants live in constant memory so that the DFMAs take them as
// constant-bank operands instead of re-materialising 64-bit immediates in registers
__constant__ double c_expc[8] = {
    369.32993046757464,       // 0: 256/ln2
    0.0027076061737716372,    // 1: ln2/256 high part (low 20 mantissa bits zero)
    2.9064910585985925e-13,   // 2: ln2/256 low part
    4.16666666666666666667e-02, 1.66666666666666666667e-01, 0.5, 1.0, 1.0};  // 3..7: 1/4! .. 1/0!

// Largest |x| for which the unchecked variant is valid (k = x*32/ln2 must fit 32 bits).
#define MEDGP_EXP_UNCHECKED_MAX 5.0e6

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
    for (int c = 4; c < 8; c++)
#pragma unroll
        for (int i = 0; i < NV; i++) p[i] = fma(p[i], r[i], c_expc[c]);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const double tj = s_tab[k[i] & (MEDGP_EXP_TAB - 1)];
        int m = k[i] >> 8;
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
