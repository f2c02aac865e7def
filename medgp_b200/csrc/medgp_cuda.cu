// medgp_cuda.cu -- C ABI (include/medgp_cuda.h) and host orchestration of libmedgp_cuda.so.
//
// One context = one GPU + one stream + one workspace arena.  A call receives a batch of
// evaluations (series id, theta); they are sorted by padded size, cut into chunks that fit
// the arena, and each chunk runs the stage sequence
//   prep -> assemble -> potrf (diag + panel per block column) -> [cross] -> solve
//        -> [trtri rows -> alpha -> lauum -> grad -> grad_finish]      (want_grad)
//        -> [pred_finish]                                              (prediction)
// with grid.y (or grid.x) indexing the evaluations of the chunk.  No CPU fallback exists:
// without a device medgp_cuda_create fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <string>
#include <vector>

#include "../../include/medgp_cuda.h"
#include "common.cuh"
#include "cov.cuh"
#include "linalg.cuh"
#include "scg.cuh"
#include "kde.cuh"

#define MEDGP_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr int kGradRows = 32;   // rows per gradient work item (= one warp, lane per row)
constexpr int kGradCols = 64;   // target columns per gradient work item
constexpr int kMaxJitter = 10;  // inference/c_inference_exact.cpp:99

// device memory shared by the series of one upload; when the last of them goes it returns to
// the context's small pool (test-time workloads upload and drop a batch of windows per super-step:
// cudaMalloc / cudaFree, which synchronise the device, stay out of that loop)
struct BlobPool {
    std::vector<std::pair<char *, size_t> > free_list;
    static constexpr size_t kMaxPooled = 8;
    char *take(size_t bytes, size_t &cap)
    {
        size_t best = free_list.size();
        for (size_t i = 0; i < free_list.size(); i++)
            if (free_list[i].second >= bytes && (best == free_list.size() || free_list[i].second < free_list[best].second)) best = i;
        if (best < free_list.size() && free_list[best].second <= 4 * bytes + (1 << 20)) {
            char *p = free_list[best].first;
            cap = free_list[best].second;
            free_list.erase(free_list.begin() + best);
            return p;
        }
        char *p = nullptr;
        cap = bytes + bytes / 4;
        if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr;
        return p;
    }
    void give(char *p, size_t cap)
    {
        if (free_list.size() >= kMaxPooled) {  // drop the smallest
            size_t small = 0;
            for (size_t i = 1; i < free_list.size(); i++)
                if (free_list[i].second < free_list[small].second) small = i;
            if (free_list[small].second < cap) {
                cudaFree(free_list[small].first);
                free_list[small] = std::make_pair(p, cap);
            } else {
                cudaFree(p);
            }
            return;
        }
        free_list.push_back(std::make_pair(p, cap));
    }
    void clear()
    {
        for (auto &e : free_list) cudaFree(e.first);
        free_list.clear();
    }
};

struct DeviceBlob {
    char *d = nullptr;
    size_t cap = 0;
    BlobPool *pool = nullptr;
    DeviceBlob(char *p, size_t c, BlobPool *pl) : d(p), cap(c), pool(pl) {}
    ~DeviceBlob() { if (d) { if (pool) pool->give(d, cap); else cudaFree(d); } }
    DeviceBlob(const DeviceBlob &) = delete;
    DeviceBlob &operator=(const DeviceBlob &) = delete;
};

struct Series {
    bool alive = false;
    std::shared_ptr<DeviceBlob> blob;
    int n = 0, npad = 0, T = 0, nitems = 0, nseg = 0;
    double trange2 = 0.0;   // (max t - min t)^2
    bool time_order = false; // points sorted by time (online imputation) instead of by feature
    bool given_order = false; // points kept in the caller's order (factor export)
    bool grad_ok = true;      // feature-major: the gradient kernel's work items exist
    int ngroups = 0;         // same-timestamp groups of a time-ordered series
    int *d_gstart = nullptr, *d_perm = nullptr;
    int2 *d_frange = nullptr;  // feature range of every 64-point block
    double *d_t = nullptr, *d_y = nullptr;
    int *d_meta = nullptr, *d_off = nullptr, *d_seg_start = nullptr;
    int4 *d_items = nullptr;
    std::vector<int> perm;  // internal position -> caller position
};

struct Request {
    int series, out_index, jitter, nstar, star_off;
};

struct StageMark {
    int stage;
    cudaEvent_t a, b;
    int stream;  // sub-stream index, -1 = the context's stream (timeline dumps only)
};

struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    long long launches[MEDGP_STAGE_COUNT] = {};
    uint64_t last_use = 0;  // LRU stamp
};

constexpr int kDescSlots = 4;      // ring of page-locked descriptor staging buffers
constexpr int kTicketsPerSub = 256; // role counters of k_potrf_step, one per block column
constexpr size_t kGraphCacheMax = 128;

}  // namespace

struct medgp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool model_set = false;
    ModelDims md{};
    char *arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<Series> series;
    std::vector<int> free_slots;  // dead entries of `series`
    BlobPool blob_pool;           // device memory of dropped uploads, reused by the next ones
    char *h_upload = nullptr;     // page-locked staging of an upload
    size_t upload_cap = 0;
    std::string err;
    // staging owned by the context (grown on demand)
    // h_descs is the current slot of a ring: a call fills its own slot, so the next call need
    // not wait for this call's descriptor upload (no host synchronisation on the device path)
    EvalDesc *h_descs = nullptr, *d_descs = nullptr;
    EvalDesc *h_desc_ring[kDescSlots] = {};
    cudaEvent_t desc_ev[kDescSlots] = {};
    bool desc_ev_pending[kDescSlots] = {};
    int desc_slot = 0;
    size_t desc_cap = 0;
    int *d_tickets = nullptr;  // 8 sub-chunk streams x kTicketsPerSub
    bool device_retry = true;  // MEDGP_DEVICE_RETRY=0: jitter retries are driven by the host instead of the graph's WHILE node
    const int *ext_skip = nullptr;  // device flags by out_index (optimiser sessions): 1 = pass over this evaluation
    bool retry_on_device = false;  // set by run_batch: the last chunk sequence carried its own jitter loop
    int force_fail = 0;        // medgp_cuda_debug_force_fail: the first attempts of every evaluation are declared failed
    uint64_t graph_clock = 0;
    double *h_theta = nullptr, *d_theta = nullptr, *h_out = nullptr, *d_out = nullptr;
    size_t theta_cap = 0, out_cap = 0;
    int *h_status = nullptr, *d_status = nullptr, *d_fail = nullptr;
    size_t status_cap = 0, fail_cap = 0;
    double *d_star_t = nullptr;
    int *d_star_meta = nullptr;
    size_t star_cap = 0;
    // sub-chunk streams (fork/join around the context's stream)
    int max_streams = 1;
    bool use_graphs = true;  // MEDGP_GRAPHS=0 disables CUDA-graph replay of chunk launch sequences
    // device + pinned blocks of finished optimiser sessions, kept for the next session: cudaFree /
    // cudaMalloc were observed to take up to 0.7 s each on a busy context, and a new session at the
    // old addresses also finds its launch sequences in the graph cache
    struct ScgBlock { char *dev = nullptr; size_t dev_bytes = 0; int *host = nullptr; size_t host_bytes = 0; bool busy = false; };
    std::vector<ScgBlock> scg_blocks;
    bool allow_direct = false;          // set by the host-buffer entry points around run_batch (capture on second sighting)
    bool lazy_capture = true;           // MEDGP_LAZY_CAPTURE=0: capture at first sighting everywhere
    std::set<uint64_t> seen_keys;       // chunk structures issued directly once
    std::map<uint64_t, GraphEntry> graphs;
    bool fuse_diag = true;   // MEDGP_FUSE_DIAG=0: separate diagonal kernels in the left-looking path
    bool chain_diag = false; // MEDGP_CHAIN_DIAG=1: diagonal blocks k >= 1 are factored inside the panel kernel of step k-1
    int gemm_smem_pad = 0;  // MEDGP_GEMM_SMEM_PAD: extra dynamic smem of the tile-GEMM kernels (lowers their CTAs/SM; experiments)
    int stagger_us = 0;     // MEDGP_STAGGER_US: start sub-chunk stream s that many microseconds x s late
    int fold_max = 128;     // MEDGP_FOLD_MAX: chunks with fewer matrices fold panel tiles into the diagonal blocks
    int force_rl = -1;  // MEDGP_RL=0/1 forces the left-/right-looking factorisation (experiments)
    cudaStream_t sub_streams[8] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[8] = {};
    // look-ahead of the right-looking factorisation: a second stream per sub-chunk for the bulk of
    // the trailing update, which runs beside the next step's diagonal block and panel
    cudaStream_t aux_streams[8] = {};
    cudaEvent_t ev_panel[8] = {}, ev_bulk[8] = {};
    bool lookahead = true;  // MEDGP_LOOKAHEAD=0 disables it
    int flow = -1;          // MEDGP_FLOW bits: 1 = dataflow factorisation, 2 = dataflow triangular inverse (one launch each,
                            // per-tile flags); -1 = chosen per chunk
    size_t rl_count_now = 0; // evaluations in the chunk being issued
    int rl_width = 0;       // block columns per panel of the right-looking factorisation (MEDGP_RL_W); 0 = by the number
                            // of matrices in the chunk: 2 for one or two (shortest critical path), 4 beyond (measured at n = 4000)
    int rl_width_now = 2;   // the width of the chunk being issued
    cudaEvent_t ev_t0 = nullptr;      // MEDGP_TIMELINE: origin of the dumped stage intervals
    const char *timeline = nullptr;   // MEDGP_TIMELINE=<file>: with profiling on, keep the sub-streams and dump every stage interval
    // profiling
    bool profile = false;
    std::vector<StageMark> marks, open_marks;
    std::vector<cudaEvent_t> event_pool;
    medgp_stage_times times{};
};

namespace {

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                    \
            return MEDGP_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void fill_dims(ModelDims &md, int Q, int D, int R, double pi)
{
    md.Q = Q; md.D = D; md.R = R; md.pi = pi;
    md.P = D + Q * (D * R + 2 + D);
    int o = 0;
    md.oB = o; o += Q * D * D;
    md.oSig2 = o; o += D;
    md.oW = o; o += Q;
    md.oC = o; o += Q;
    md.oA = o; o += Q * D * R;
    md.oKappa = o; o += Q * D;
    md.oBdiag = o; o += D;
    md.parLen = (o + 1) & ~1;
}

// bytes of arena one evaluation needs
size_t eval_bytes(const ModelDims &md, const Series &s, int nrhs, bool grad)
{
    const size_t np = s.npad;
    size_t b = 0;
    b += align_up((size_t)s.T * s.T * kTileElems * 8, 256);   // M (tile-major)
    b += 2 * align_up((size_t)s.T * kTileElems * 8, 256);     // dinv, dinvT
    b += align_up(np * (size_t)nrhs * 8, 256);             // rhs
    b += align_up(np * 8, 256);                            // alpha
    b += align_up(np * (size_t)md.Q * 16, 256);            // cs
    b += align_up((size_t)md.parLen * 8, 256);             // par
    b += align_up((size_t)s.T * 8, 256);                   // blk
    b += align_up((size_t)s.T * s.T * 4, 256);             // flags (one per tile)
    if (grad) b += align_up((size_t)s.nseg * md.D * (3 * md.Q + 1) * 8, 256);
    return b;
}

// Grows the context-owned staging buffers.  Growing means freeing buffers the stream may still
// be using, so the stream is drained first; the steady state (same batch size as before) does
// not synchronise.
int ensure_staging(medgp_ctx *ctx, size_t nreq, size_t nstar)
{
    const size_t P = ctx->md.P;
    const size_t need_out = nreq * (P + 1) + 2 * nstar;  // nlml (nreq) + grad (nreq*P) + mean/var (2*nstar)
    if (nreq > ctx->desc_cap || nreq * P > ctx->theta_cap || need_out > ctx->out_cap || nreq > ctx->status_cap ||
        nreq > ctx->fail_cap || nstar > ctx->star_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < kDescSlots; i++) ctx->desc_ev_pending[i] = false;
    }
    if (nreq > ctx->desc_cap) {
        for (int i = 0; i < kDescSlots; i++)
            if (ctx->h_desc_ring[i]) cudaFreeHost(ctx->h_desc_ring[i]);
        if (ctx->d_descs) cudaFree(ctx->d_descs);
        ctx->desc_cap = std::max(nreq, 2 * ctx->desc_cap);
        for (int i = 0; i < kDescSlots; i++) CU(cudaMallocHost(&ctx->h_desc_ring[i], ctx->desc_cap * sizeof(EvalDesc)));
        CU(cudaMalloc(&ctx->d_descs, ctx->desc_cap * sizeof(EvalDesc)));
        // the descriptor array is baked into the captured launch sequences
        for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second.exec);
        ctx->graphs.clear();
    }
    if (nreq * P > ctx->theta_cap) {
        if (ctx->h_theta) cudaFreeHost(ctx->h_theta);
        if (ctx->d_theta) cudaFree(ctx->d_theta);
        ctx->theta_cap = std::max(nreq * P, 2 * ctx->theta_cap);
        CU(cudaMallocHost(&ctx->h_theta, ctx->theta_cap * 8));
        CU(cudaMalloc(&ctx->d_theta, ctx->theta_cap * 8));
    }
    if (need_out > ctx->out_cap) {
        if (ctx->h_out) cudaFreeHost(ctx->h_out);
        if (ctx->d_out) cudaFree(ctx->d_out);
        ctx->out_cap = std::max(need_out, 2 * ctx->out_cap);
        CU(cudaMallocHost(&ctx->h_out, ctx->out_cap * 8));
        CU(cudaMalloc(&ctx->d_out, ctx->out_cap * 8));
    }
    if (nreq > ctx->status_cap) {
        if (ctx->h_status) cudaFreeHost(ctx->h_status);
        if (ctx->d_status) cudaFree(ctx->d_status);
        ctx->status_cap = std::max(nreq, 2 * ctx->status_cap);
        CU(cudaMallocHost(&ctx->h_status, ctx->status_cap * sizeof(int)));
        CU(cudaMalloc(&ctx->d_status, ctx->status_cap * sizeof(int)));
    }
    if (nreq > ctx->fail_cap) {
        if (ctx->d_fail) cudaFree(ctx->d_fail);
        ctx->fail_cap = std::max(nreq, 2 * ctx->fail_cap);
        CU(cudaMalloc(&ctx->d_fail, ctx->fail_cap * sizeof(int)));
    }
    if (nstar > ctx->star_cap) {
        if (ctx->d_star_t) cudaFree(ctx->d_star_t);
        if (ctx->d_star_meta) cudaFree(ctx->d_star_meta);
        ctx->star_cap = std::max(nstar, 2 * ctx->star_cap);
        CU(cudaMalloc(&ctx->d_star_t, ctx->star_cap * 8));
        CU(cudaMalloc(&ctx->d_star_meta, ctx->star_cap * sizeof(int)));
    }
    // this call's descriptor slot: wait until the upload that last used it has been consumed
    ctx->desc_slot = (ctx->desc_slot + 1) % kDescSlots;
    if (ctx->desc_ev_pending[ctx->desc_slot]) {
        CU(cudaEventSynchronize(ctx->desc_ev[ctx->desc_slot]));
        ctx->desc_ev_pending[ctx->desc_slot] = false;
    }
    ctx->h_descs = ctx->h_desc_ring[ctx->desc_slot];
    return MEDGP_OK;
}

// after the last descriptor upload of a call has been enqueued
int release_desc_slot(medgp_ctx *ctx)
{
    CU(cudaEventRecord(ctx->desc_ev[ctx->desc_slot], ctx->stream));
    ctx->desc_ev_pending[ctx->desc_slot] = true;
    return MEDGP_OK;
}

cudaEvent_t get_event(medgp_ctx *ctx)
{
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

// per-stage CUDA events (profiling only): begin pushes the start event, end closes the mark
void stage_begin(medgp_ctx *ctx, int stage, cudaStream_t st)
{
    if (!ctx->profile) return;
    cudaEvent_t a = get_event(ctx);
    cudaEventRecord(a, st);
    int sidx = -1;
    for (int i = 0; i < 8; i++)
        if (ctx->sub_streams[i] == st) sidx = i;
    ctx->open_marks.push_back({stage, a, nullptr, sidx});
}

void stage_end(medgp_ctx *ctx, int stage, cudaStream_t st)
{
    if (!ctx->profile) return;
    for (size_t i = ctx->open_marks.size(); i-- > 0;)
        if (ctx->open_marks[i].stage == stage && ctx->open_marks[i].stream == [&]() {
                int sidx = -1;
                for (int q = 0; q < 8; q++)
                    if (ctx->sub_streams[q] == st) sidx = q;
                return sidx; }()) {
            StageMark m = ctx->open_marks[i];
            ctx->open_marks.erase(ctx->open_marks.begin() + i);
            m.b = get_event(ctx);
            cudaEventRecord(m.b, st);
            ctx->marks.push_back(m);
            return;
        }
}

void resolve_marks(medgp_ctx *ctx)
{
    FILE *tl = (ctx->timeline && ctx->ev_t0 && !ctx->marks.empty()) ? fopen(ctx->timeline, "a") : nullptr;
    for (auto &m : ctx->marks) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, m.a, m.b) == cudaSuccess) ctx->times.ms[m.stage] += ms;
        if (tl) {
            float t_a = 0.f;
            cudaEventElapsedTime(&t_a, ctx->ev_t0, m.a);
            fprintf(tl, "%d %d %.4f %.4f\n", m.stream, m.stage, t_a, t_a + ms);
        }
        ctx->event_pool.push_back(m.a);
        ctx->event_pool.push_back(m.b);
    }
    if (tl) {
        fprintf(tl, "# end of call\n");
        fclose(tl);
    }
    ctx->marks.clear();
}

// One sub-chunk = a contiguous descriptor range launched on one stream.
struct SubChunk {
    size_t base = 0, cnt = 0;
    int Tmax = 0, items_max = 0, nstar_max = 0, groups_max = 0;
    std::vector<int> T;  // per evaluation, descending
    unsigned act(int k) const  // evaluations with T > k (a prefix: sorted descending)
    {
        size_t lo = 0, hi = cnt;
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            if (T[mid] > k) lo = mid + 1; else hi = mid;
        }
        return (unsigned)lo;
    }
};

template <int QT>
void launch_grad_q(dim3 gg, cudaStream_t st, const EvalDesc *dd, const ModelDims &md)
{
    k_grad<QT><<<gg, 32 * MEDGP_GW, grad_smem_bytes<QT>(), st>>>(dd, md);
}

template <int QT>
void launch_assemble_q(dim3 gg, int smem, cudaStream_t st, const EvalDesc *dd, const ModelDims &md)
{
    k_assemble<QT><<<gg, 256, smem, st>>>(dd, md);
}

void launch_assemble(int Q, dim3 gg, int smem, cudaStream_t st, const EvalDesc *dd, const ModelDims &md)
{
    switch (Q) {
        case 1: launch_assemble_q<1>(gg, smem, st, dd, md); break;
        case 2: launch_assemble_q<2>(gg, smem, st, dd, md); break;
        case 3: launch_assemble_q<3>(gg, smem, st, dd, md); break;
        case 4: launch_assemble_q<4>(gg, smem, st, dd, md); break;
        case 5: launch_assemble_q<5>(gg, smem, st, dd, md); break;
        case 6: launch_assemble_q<6>(gg, smem, st, dd, md); break;
        case 7: launch_assemble_q<7>(gg, smem, st, dd, md); break;
        default: launch_assemble_q<8>(gg, smem, st, dd, md); break;
    }
}

template <int QT>
void set_assemble_smem_q(int bytes)
{
    cudaFuncSetAttribute(k_assemble<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_grad<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, grad_smem_bytes<QT>());
}

void launch_grad(int Q, dim3 gg, cudaStream_t st, const EvalDesc *dd, const ModelDims &md)
{
    switch (Q) {
        case 1: launch_grad_q<1>(gg, st, dd, md); break;
        case 2: launch_grad_q<2>(gg, st, dd, md); break;
        case 3: launch_grad_q<3>(gg, st, dd, md); break;
        case 4: launch_grad_q<4>(gg, st, dd, md); break;
        case 5: launch_grad_q<5>(gg, st, dd, md); break;
        case 6: launch_grad_q<6>(gg, st, dd, md); break;
        case 7: launch_grad_q<7>(gg, st, dd, md); break;
        default: launch_grad_q<8>(gg, st, dd, md); break;
    }
}

// The stage sequence of one sub-chunk on stream st, as a list of launch closures so that the
// caller can interleave the sub-chunks' launches round-robin (every stream starts at once
// instead of waiting for the CPU to issue all launches of the streams before it).
typedef std::vector<std::function<void()> > LaunchList;

void build_sub(medgp_ctx *ctx, const SubChunk &sc, bool rl, bool fold, cudaStream_t st, const double *d_theta, int mode,
               double *d_nlml, double *d_grad, int *d_status, double *d_mean, double *d_var, LaunchList &out,
               int stagger_slot = 0)
{
    const ModelDims md = ctx->md;
    const bool grad = (mode == 1), pred = (mode == 2);
    const int gemm_smem = kGemmSmemBytes + ctx->gemm_smem_pad;
    const int asm_smem = (md.Q * md.D * md.D + md.Q) * 8;
    const int npairs = md.D * (md.D + 1) / 2;
    const int fin_smem = (md.Q * md.D * md.D + 2 * npairs * md.Q + md.D) * 8;
    const EvalDesc *dd = ctx->d_descs + sc.base;
    const unsigned ncta = (unsigned)sc.cnt;
    const int Tmax = sc.Tmax, ntri = Tmax * (Tmax + 1) / 2;
    long long *L = ctx->times.launches;
    int *d_fail = ctx->d_fail;
    const SubChunk *scp = &sc;
    int *tickets = ctx->d_tickets + (size_t)stagger_slot * kTicketsPerSub;  // stagger_slot = sub-chunk index
    const int force_fail = ctx->force_fail;
    // stage markers (profiling): begin/end closures record events on the stream
    auto begin = [&](int stage) { out.push_back([=]() { stage_begin(ctx, stage, st); }); };
    auto end = [&](int stage) { out.push_back([=]() { stage_end(ctx, stage, st); }); };

    if (stagger_slot > 0 && ctx->stagger_us > 0) {
        const unsigned long long ns = 1000ULL * (unsigned long long)ctx->stagger_us * (unsigned long long)stagger_slot;
        out.push_back([=]() { k_delay<<<1, 1, 0, st>>>(ns); });
    }
    begin(MEDGP_STAGE_PREP);
    out.push_back([=]() { k_prep<<<ncta, 256, 0, st>>>(dd, md, d_theta, tickets, kTicketsPerSub); L[MEDGP_STAGE_PREP]++; });
    end(MEDGP_STAGE_PREP);
    begin(MEDGP_STAGE_ASSEMBLE);
    out.push_back([=]() { launch_assemble(md.Q, dim3(ntri, ncta), asm_smem, st, dd, md); L[MEDGP_STAGE_ASSEMBLE]++; });
    end(MEDGP_STAGE_ASSEMBLE);
    if (pred && sc.nstar_max > 0) {  // cross-covariance columns ride along the factorisation
        const int nsm = sc.nstar_max;
        begin(MEDGP_STAGE_PREDICT);
        out.push_back([=]() { k_cross<<<dim3(nsm, ncta), 256, 0, st>>>(dd, md); L[MEDGP_STAGE_PREDICT]++; });
        end(MEDGP_STAGE_PREDICT);
    }
    // left-looking with few matrices in the chunk: ONE launch per block column (k_potrf_step: the
    // diagonal CTA factors block k while the panel CTAs run their k-tile products, then wait on
    // its flag).  Large batches keep separate diagonal / panel kernels: their sub-chunk streams
    // already overlap, and CTAs spinning on a flag would hold slots the other streams can use.
    const bool step_kernel = !rl && fold && ctx->fuse_diag && Tmax <= kTicketsPerSub;  // fold <=> few matrices in the chunk
    const bool look = rl && ctx->lookahead;
    cudaStream_t st2 = ctx->aux_streams[stagger_slot];
    cudaEvent_t ev_panel = ctx->ev_panel[stagger_slot], ev_bulk = ctx->ev_bulk[stagger_slot];
    bool bulk_pending = false;
    // Panel-blocked right-looking schedule (few large matrices): W block columns are factored
    // left-looking among themselves, then the trailing matrix gets ONE W-tile-deep update.
    // Look-ahead: the block columns of the next panel are updated on this stream (the critical
    // path: next diagonal blocks and panels need them), everything right of them on the auxiliary
    // stream beside the next panel's factorisation.  Orderings that matter: the bulk update of a
    // panel follows its last panel kernel; the next-panel part of panel p follows the bulk of
    // panel p-1 (both touch the same block columns); bulks are ordered by their stream.
    // Dataflow path: ONE launch for the factorisation, ONE for the triangular inverse (linalg.cuh)
    // Measured (n = 4000 with 1 / 5 / 32 matrices in flight, C2, C3, batch-1 latencies): the dataflow inverse
    // wins for few large matrices (-10 % / -7 % on NLML+gradient with 1 / 5 in flight) and loses 1-3 % for
    // many small ones; the dataflow factorisation wins only where little is in flight (its resident roles
    // advance one tile product per finished block column, so it is latency-, not throughput-oriented).
    // Small chunks (one or two matrices of any size, or a handful of small ones: the batch-1 shape of
    // main_one_train's line searches) take both: n = 300 / 500 with 1-5 evaluations per call -20 % latency.
    // (tools/sweep_small.py, tools/sweep_mid.py: up to 16 block rows the dataflow pair wins by 15-30 % until
    // count x block rows reaches about 700 -- 48 matrices of n = 900 -- and loses beyond; with more block rows
    // it only wins for one or two matrices, the right-looking schedule takes over from three)
    const size_t cnt_now = ctx->rl_count_now;
    const bool small_chunk = cnt_now <= 2 || cnt_now * (size_t)Tmax <= 48 || (Tmax <= 16 && cnt_now * (size_t)Tmax <= 720);
    const int flow_bits = ctx->flow >= 0 ? ctx->flow : (small_chunk ? 3 : (rl ? 2 : 0));
    const bool flow = (flow_bits & 1) != 0 && Tmax <= MEDGP_FLOW_TMAX;        // bit 0: factorisation
    const bool flow_trtri = (flow_bits & 2) != 0 && Tmax <= MEDGP_FLOW_TMAX;  // bit 1: triangular inverse
    FlowMap fm_potrf{}, fm_trtri{};
    // both: the inverse's roles ride in the factorisation's launch (k_potrf_flow with_inverse)
    const bool flow_both = flow && flow_trtri && (grad || mode == 4);
    if (flow || flow_trtri) {
        fm_potrf.Tmax = fm_trtri.Tmax = Tmax;
        for (int t = 0; t <= Tmax; t++) fm_potrf.act[t] = fm_trtri.act[t] = (int)sc.act(t);
        int tot = 0;
        for (int c = 0; c < Tmax; c++) {  // block column c (roles: see k_potrf_flow)
            fm_potrf.base[c] = tot;
            tot += c == 0 ? fm_potrf.act[0] : (c + 1 < Tmax ? fm_potrf.act[c + 1] : 0);  // DIAG0 / PRE(c+1)
            for (int r = 1; c + r < Tmax; r++) tot += fm_potrf.act[c + r];               // PANEL(c + r, c)
            if (flow_both) tot += c * fm_potrf.act[c];                                   // INV(j, c), j < c
        }
        fm_potrf.base[Tmax] = fm_potrf.total = tot;
        tot = 0;
        fm_trtri.base[0] = 0;
        for (int i = 1; i < Tmax; i++) {  // block row i of L^-1: i roles per evaluation that has it
            fm_trtri.base[i] = tot;
            tot += i * fm_trtri.act[i];
        }
        fm_trtri.base[Tmax] = fm_trtri.total = tot;
    }
    // (ticket counters: the last two of the sub-chunk's block; the first Tmax belong to k_potrf_step)
    if (flow) {
        const FlowMap fm = fm_potrf;
        begin(MEDGP_STAGE_POTRF);
        out.push_back([=]() { k_potrf_flow<<<fm.total, MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, fm, d_fail, tickets + kTicketsPerSub - 2, flow_both ? 1 : 0); L[MEDGP_STAGE_POTRF]++; });
        end(MEDGP_STAGE_POTRF);
    }
    if (rl && !flow) {
        const int W = std::max(1, ctx->rl_width_now);
        const bool rl_step = ctx->fuse_diag && Tmax <= kTicketsPerSub;
        for (int k0 = 0; k0 < Tmax; k0 += W) {
            const int k1 = std::min(k0 + W, Tmax);
            for (int k = k0; k < k1; k++) {
                const unsigned a0 = sc.act(k), a1 = k + 1 < Tmax ? sc.act(k + 1) : 0;
                const int rem = Tmax - k - 1, depth = k - k0;
                if (rl_step) {  // one launch per block column: the panel roles wait for the diagonal role's flag
                    begin(MEDGP_STAGE_POTRF);
                    out.push_back([=]() { k_potrf_step<<<dim3(a0, rem + 1), MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, k, d_fail, tickets + k, k0, 0); L[MEDGP_STAGE_POTRF]++; });
                    end(MEDGP_STAGE_POTRF);
                    continue;
                }
                begin(MEDGP_STAGE_DIAG);
                out.push_back([=]() { k_potrf_diag<<<a0, MEDGP_DIAG_THREADS, gemm_smem, st>>>(dd, k, depth, d_fail, k0); L[MEDGP_STAGE_DIAG]++; });
                end(MEDGP_STAGE_DIAG);
                if (rem > 0 && a1 > 0) {
                    begin(MEDGP_STAGE_POTRF);
                    out.push_back([=]() { k_potrf_panel<<<dim3(rem, a1), MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, k, depth, 0, d_fail, k0); L[MEDGP_STAGE_POTRF]++; });
                    end(MEDGP_STAGE_POTRF);
                }
            }
            const unsigned at = k1 < Tmax ? sc.act(k1) : 0;
            if (at == 0) continue;
            const int nk = k1 - k0, ncolA = std::min(W, Tmax - k1), jB = k1 + ncolA, remB = Tmax - jB;
            begin(MEDGP_STAGE_POTRF);
            if (!look) {
                const int rem = Tmax - k1;
                out.push_back([=]() { k_syrk_update<<<dim3(rem * (rem + 1) / 2, at), MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, k0, nk, k1, 0); L[MEDGP_STAGE_POTRF]++; });
            } else {
                if (remB > 0) {
                    out.push_back([=]() {
                        cudaEventRecord(ev_panel, st);
                        cudaStreamWaitEvent(st2, ev_panel, 0);
                        k_syrk_update<<<dim3(remB * (remB + 1) / 2, at), MEDGP_GEMM_THREADS, gemm_smem, st2>>>(dd, k0, nk, jB, 0);
                        L[MEDGP_STAGE_POTRF]++;
                    });
                }
                if (bulk_pending) out.push_back([=]() { cudaStreamWaitEvent(st, ev_bulk, 0); });
                out.push_back([=]() { k_syrk_update<<<dim3(ncolA * (Tmax - k1), at), MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, k0, nk, k1, ncolA); L[MEDGP_STAGE_POTRF]++; });
                if (remB > 0) {
                    out.push_back([=]() { cudaEventRecord(ev_bulk, st2); });
                    bulk_pending = true;
                } else {
                    bulk_pending = false;  // the wait above joined the last bulk update
                }
            }
            end(MEDGP_STAGE_POTRF);
        }
    }
    for (int k = 0; k < Tmax && !rl && !flow; k++) {
        const unsigned a0 = sc.act(k), a1 = k + 1 < Tmax ? sc.act(k + 1) : 0;
        const int rem = Tmax - k - 1;
        if (step_kernel) {
            begin(MEDGP_STAGE_POTRF);
            out.push_back([=]() { k_potrf_step<<<dim3(a0, rem + 1), MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, k, d_fail, tickets + k); L[MEDGP_STAGE_POTRF]++; });
            end(MEDGP_STAGE_POTRF);
            continue;
        }
        // chain_diag: the panel CTA of row k+1 folds its tile into K_{k+1,k+1} and factors that
        // block on the spot, so only block 0 needs a diagonal launch and the 64-pivot chain of
        // block k+1 hides behind the other panel CTAs of step k
        const bool chain_diag = ctx->chain_diag;
        const int depth = (fold || chain_diag) ? 0 : k;
        const int pdepth = k, pfold = chain_diag ? 2 : (fold ? 1 : 0);
        if (!chain_diag || k == 0) {
            begin(MEDGP_STAGE_DIAG);
            out.push_back([=]() { k_potrf_diag<<<a0, MEDGP_DIAG_THREADS, gemm_smem, st>>>(dd, k, depth, d_fail, 0); L[MEDGP_STAGE_DIAG]++; });
            end(MEDGP_STAGE_DIAG);
        }
        if (rem > 0) {
            begin(MEDGP_STAGE_POTRF);
            out.push_back([=]() { k_potrf_panel<<<dim3(rem, a1), MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, k, pdepth, pfold, d_fail, 0); L[MEDGP_STAGE_POTRF]++; });
            end(MEDGP_STAGE_POTRF);
        }
    }
    if (bulk_pending) out.push_back([=]() { cudaStreamWaitEvent(st, ev_bulk, 0); });  // join the auxiliary stream
    begin(MEDGP_STAGE_SOLVE);
    out.push_back([=]() { k_solve<<<ncta, 256, 0, st>>>(dd, md, d_nlml, d_status, d_fail, force_fail); L[MEDGP_STAGE_SOLVE]++; });
    end(MEDGP_STAGE_SOLVE);
    if (grad || mode == 4) {
        begin(MEDGP_STAGE_TRTRI);
        if (flow_trtri && !flow_both && fm_trtri.total > 0) {
            const FlowMap fm = fm_trtri;
            out.push_back([=]() { k_trtri_flow<<<fm.total, MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, fm, tickets + kTicketsPerSub - 1); L[MEDGP_STAGE_TRTRI]++; });
        }
        for (int i = 1; i < Tmax && !flow_trtri; i++) {
            const unsigned ai = sc.act(i);
            if (rl) {
                const int k = i - 1;  // rows 0..k of U are final: push them into rows i > k
                out.push_back([=]() { k_trtri_update<<<dim3((Tmax - k - 1) * (k + 1), ai), MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, k); L[MEDGP_STAGE_TRTRI]++; });
            }
            out.push_back([=]() { k_trtri_row<<<dim3(i, ai), MEDGP_GEMM_THREADS, gemm_smem, st>>>(dd, i, rl ? 1 : 0); L[MEDGP_STAGE_TRTRI]++; });
        }
        end(MEDGP_STAGE_TRTRI);
    }
    if (mode == 4) {  // factor export: alpha from U, no K^-1
        begin(MEDGP_STAGE_SOLVE);
        out.push_back([=]() { k_alpha<<<dim3(Tmax, ncta), 128, 0, st>>>(dd); L[MEDGP_STAGE_SOLVE]++; });
        end(MEDGP_STAGE_SOLVE);
    }
    if (grad) {
        begin(MEDGP_STAGE_LAUUM);
        out.push_back([=]() { k_lauum<<<dim3(ntri, ncta), MEDGP_GEMM_THREADS, kLauumSmemBytes + ctx->gemm_smem_pad, st>>>(dd); L[MEDGP_STAGE_LAUUM]++; });
        end(MEDGP_STAGE_LAUUM);
        begin(MEDGP_STAGE_GRAD);
        const int items = scp->items_max;
        out.push_back([=]() {
            launch_grad(md.Q, dim3((items + MEDGP_GW - 1) / MEDGP_GW, ncta), st, dd, md);
            k_grad_finish<<<ncta, 256, fin_smem, st>>>(dd, md, d_grad, d_fail);
            L[MEDGP_STAGE_GRAD] += 2;
        });
        end(MEDGP_STAGE_GRAD);
    }
    if (pred && sc.nstar_max > 0) {
        const int nsm = sc.nstar_max;
        begin(MEDGP_STAGE_PREDICT);
        out.push_back([=]() { k_pred_finish<<<dim3(nsm, ncta), 256, 0, st>>>(dd, md, d_mean, d_var, d_fail); L[MEDGP_STAGE_PREDICT]++; });
        end(MEDGP_STAGE_PREDICT);
    }
    if (mode == 3 && sc.groups_max > 0) {
        const int ng = sc.groups_max;
        begin(MEDGP_STAGE_PREDICT);
        out.push_back([=]() { k_online<<<dim3(ng, ncta), 32, 0, st>>>(dd, d_mean, d_var, d_fail); L[MEDGP_STAGE_PREDICT]++; });
        end(MEDGP_STAGE_PREDICT);
    }
}

// The core: run `reqs` (any sizes) through the stage sequence.  d_theta is indexed by
// out_index.  mode: 0 = NLML only, 1 = NLML + gradient, 2 = prediction, 3 = online imputation
// (time-ordered series; star_off of the request = offset of the series in d_mean / d_var),
// 4 = NLML + triangular inverse + alpha (medgp_cuda_export_factors).
// Evaluations are sorted by size, cut into chunks that fit the arena, and every chunk is dealt
// round-robin into up to kMaxStreams sub-chunks that run on their own streams, so the
// latency-bound phases of one sub-chunk (diagonal blocks, small trtri rows, tails) overlap the
// tensor-core phases of the others.
int run_batch(medgp_ctx *ctx, std::vector<Request> reqs, const double *d_theta, int mode,
              double *d_nlml, double *d_grad, int *d_status, double *d_mean, double *d_var,
              size_t desc_base)
{
    const ModelDims &md = ctx->md;
    cudaStream_t st = ctx->stream;
    const bool grad = (mode == 1), pred = (mode == 2);
    std::stable_sort(reqs.begin(), reqs.end(), [&](const Request &a, const Request &b) {
        return ctx->series[a.series].npad > ctx->series[b.series].npad;
    });
    size_t pos = 0, dpos = desc_base;
    if (ctx->profile && ctx->timeline) {
        if (!ctx->ev_t0) cudaEventCreate(&ctx->ev_t0);
        cudaEventRecord(ctx->ev_t0, st);
    }
    while (pos < reqs.size()) {
        // ---- form a chunk
        size_t used = 0, cnt = 0;
        const size_t first = pos;
        while (pos < reqs.size() && cnt < 65535) {
            const Series &s = ctx->series[reqs[pos].series];
            const size_t b = eval_bytes(md, s, 1 + reqs[pos].nstar, grad);
            if (used + b > ctx->arena_bytes) break;
            used += b;
            pos++;
            cnt++;
        }
        if (cnt == 0) {
            ctx->err = "workspace too small for one evaluation";
            return MEDGP_ERR_NOMEM;
        }
        // ---- deal the chunk into sub-chunks (one stream each)
        const int Tbig = ctx->series[reqs[first].series].T;
        int S = 1;
        if ((!ctx->profile || ctx->timeline) && ctx->max_streams > 1) {
            S = Tbig >= 16 ? (int)std::min<size_t>(cnt, ctx->max_streams)
                           : (int)std::min<size_t>(std::max<size_t>(1, cnt / 32), ctx->max_streams);
        }
        // few large matrices -> right-looking factorisation (more CTAs per launch); many small
        // ones -> left-looking (less traffic).  Decided on the whole chunk, not per stream.
        bool rl = Tbig >= 8 && cnt * (size_t)Tbig < 600;
        if (ctx->force_rl >= 0) rl = ctx->force_rl != 0;
        ctx->rl_width_now = ctx->rl_width > 0 ? ctx->rl_width : (cnt <= 2 ? 2 : 4);
        ctx->rl_count_now = cnt;
        // left-looking with few matrices: the diagonal kernel (one CTA per matrix) must not carry
        // a k-tile product; the panel CTAs fold their tile into the diagonal block instead
        const bool fold = !rl && cnt < ctx->fold_max;
        std::vector<SubChunk> subs(S);
        // ---- carve the arena and fill descriptors, sub-chunk major
        char *p = ctx->arena;
        auto take = [&](size_t bytes) {
            char *r = p;
            p += align_up(bytes, 256);
            return r;
        };
        size_t dcur = dpos;
        // Which evaluations go to which sub-chunk.  Large chunks: CONTIGUOUS ranges of the
        // size-sorted chunk with equal shares of the n^3 work, so that the matrices of a sub-chunk
        // have (nearly) the same number of block rows: its launch grids -- sized for its largest
        // matrix -- then hold few CTAs that find nothing to do, and the CTAs of a launch are equally
        // deep.  Small chunks: round-robin (one or two matrices per stream).  MEDGP_DEAL=0: always round-robin.
        std::vector<std::vector<size_t> > members(S);
        static const bool deal_ranges = !(getenv("MEDGP_DEAL") && atoi(getenv("MEDGP_DEAL")) == 0);
        if (deal_ranges && S > 1 && cnt >= (size_t)(4 * S)) {
            std::vector<double> cum(cnt + 1, 0.0);
            for (size_t c = 0; c < cnt; c++) {
                const double t = ctx->series[reqs[first + c].series].T;
                cum[c + 1] = cum[c] + t * t * t;
            }
            size_t c = 0;
            for (int sidx = 0; sidx < S; sidx++) {
                const double upto = cum[cnt] * (double)(sidx + 1) / (double)S;
                const size_t keep = (size_t)(S - 1 - sidx);  // leave at least one evaluation for every later sub-chunk
                do members[sidx].push_back(c++);
                while (c + keep < cnt && (sidx == S - 1 || cum[c + 1] <= upto));
            }
        } else {
            for (int sidx = 0; sidx < S; sidx++)
                for (size_t c = sidx; c < cnt; c += S) members[sidx].push_back(c);
        }
        for (int sidx = 0; sidx < S; sidx++) {
            SubChunk &sc = subs[sidx];
            sc.base = dcur;
            for (size_t c : members[sidx]) {
                const Request &rq = reqs[first + c];
                const Series &s = ctx->series[rq.series];
                EvalDesc &e = ctx->h_descs[dcur++];
                const size_t np = s.npad;
                e.M = (double *)take((size_t)s.T * s.T * kTileElems * 8);
                e.dinv = (double *)take((size_t)s.T * kTileElems * 8);
                e.dinvT = (double *)take((size_t)s.T * kTileElems * 8);
                e.rhs = (double *)take(np * (size_t)(1 + rq.nstar) * 8);
                e.alpha = (double *)take(np * 8);
                e.cs = (double *)take(np * (size_t)md.Q * 16);
                e.par = (double *)take((size_t)md.parLen * 8);
                e.blk = (double *)take((size_t)s.T * 8);
                e.flags = (int *)take((size_t)s.T * s.T * 4);
                e.part = grad ? (double *)take((size_t)s.nseg * md.D * (3 * md.Q + 1) * 8) : nullptr;
                e.t = s.d_t; e.y = s.d_y; e.meta = s.d_meta; e.off = s.d_off;
                e.items = s.d_items; e.seg_start = s.d_seg_start;
                e.star_t = pred ? ctx->d_star_t + rq.star_off : nullptr;
                e.star_meta = pred ? ctx->d_star_meta + rq.star_off : nullptr;
                e.n = s.n; e.npad = s.npad; e.T = s.T; e.nitems = s.nitems;
                e.jitter = rq.jitter; e.nrhs = 1 + rq.nstar; e.nstar = rq.nstar;
                e.out_index = rq.out_index; e.star_out = rq.star_off;
                e.skip = 0; e.trange2 = s.trange2; e.pad1 = 0;
                e.gstart = s.d_gstart; e.perm = s.d_perm; e.ngroups = s.ngroups; e.frange = s.d_frange;
                sc.T.push_back(s.T);
                sc.cnt++;
                sc.Tmax = std::max(sc.Tmax, s.T);
                sc.items_max = std::max(sc.items_max, s.nitems);
                sc.nstar_max = std::max(sc.nstar_max, rq.nstar);
                sc.groups_max = std::max(sc.groups_max, s.ngroups);
                // algorithmic work
                const double n = s.n;
                ctx->times.flops[MEDGP_STAGE_POTRF] += n * n * n / 3.0;
                ctx->times.flops[MEDGP_STAGE_SOLVE] += n * n * (1 + rq.nstar);
                ctx->times.bytes[MEDGP_STAGE_ASSEMBLE] += 8.0 * n * (n + 1) / 2 + 12.0 * n;
                if (grad) {
                    ctx->times.flops[MEDGP_STAGE_TRTRI] += n * n * n / 3.0;
                    ctx->times.flops[MEDGP_STAGE_LAUUM] += n * n * n / 3.0;
                    ctx->times.bytes[MEDGP_STAGE_GRAD] += 8.0 * n * (n + 1) / 2 + 8.0 * md.P;
                }
                if (pred) ctx->times.bytes[MEDGP_STAGE_PREDICT] += rq.nstar * (8.0 * n * (n + 1) / 2 + 8.0 * n);
                ctx->times.evals++;
            }
        }
        CU(cudaMemcpyAsync(ctx->d_descs + dpos, ctx->h_descs + dpos, cnt * sizeof(EvalDesc),
                           cudaMemcpyHostToDevice, st));
        // ---- issue: a CUDA graph per chunk structure (captured once, replayed afterwards) keeps
        //      the CPU out of the way -- a step is hundreds of launches over several streams
        auto issue_all = [&]() -> int {
            if (ctx->ext_skip)  // optimiser session: retire the descriptors of finished instances first
                k_apply_skip<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(ctx->d_descs + dpos, (int)cnt, ctx->ext_skip);
            std::vector<LaunchList> prog(S);
            for (int sidx = 0; sidx < S; sidx++)
                build_sub(ctx, subs[sidx], rl, fold, S == 1 ? st : ctx->sub_streams[sidx], d_theta, mode, d_nlml,
                          d_grad, d_status, d_mean, d_var, prog[sidx], S > 1 ? sidx : 0);
            if (S > 1) {
                CU(cudaEventRecord(ctx->ev_fork, st));
                for (int sidx = 0; sidx < S; sidx++) CU(cudaStreamWaitEvent(ctx->sub_streams[sidx], ctx->ev_fork, 0));
            }
            size_t longest = 0;
            for (auto &pl : prog) longest = std::max(longest, pl.size());
            for (size_t step = 0; step < longest; step++)  // round-robin issue
                for (int sidx = 0; sidx < S; sidx++)
                    if (step < prog[sidx].size()) prog[sidx][step]();
            if (S > 1)
                for (int sidx = 0; sidx < S; sidx++) {
                    CU(cudaEventRecord(ctx->ev_join[sidx], ctx->sub_streams[sidx]));
                    CU(cudaStreamWaitEvent(st, ctx->ev_join[sidx], 0));
                }
            return MEDGP_OK;
        };
        // Jitter retries (c_inference_exact.cpp:99-108) run ON THE DEVICE: the chunk's launch
        // sequence is the body of a graph WHILE node whose last kernel (k_retry_decide) bumps the
        // jitter of the failed evaluations, retires the others, and keeps the loop alive while
        // any evaluation is left.  The common case (nothing fails) is one pass plus that kernel.
        const int max_jitter = (mode == 3) ? 0 : kMaxJitter;  // online imputation never retries (see the header)
        bool graph_mode = !ctx->profile && ctx->use_graphs;
        bool loop = graph_mode && ctx->device_retry && max_jitter > 0;
        uint64_t key = 1469598103934665603ULL;
        if (graph_mode) {
            auto mix = [&](uint64_t v) { key = (key ^ v) * 1099511628211ULL; };
            mix((uint64_t)mode); mix(rl); mix(fold); mix(ctx->fuse_diag); mix((uint64_t)ctx->stagger_us); mix(ctx->chain_diag); mix(ctx->lookahead); mix((uint64_t)(ctx->flow + 1)); mix((uint64_t)ctx->rl_width_now); mix((uint64_t)S); mix((uint64_t)dpos); mix((uint64_t)(uintptr_t)ctx->d_descs);
            mix((uint64_t)(uintptr_t)d_theta); mix((uint64_t)(uintptr_t)d_nlml); mix((uint64_t)(uintptr_t)d_grad);
            mix((uint64_t)(uintptr_t)d_status); mix((uint64_t)(uintptr_t)d_mean); mix((uint64_t)(uintptr_t)d_var);
            mix((uint64_t)(uintptr_t)ctx->d_fail); mix((uint64_t)(uintptr_t)ctx->ext_skip);
            // the launches bake in the model (ModelDims by value, Q-templated kernels, smem sizes)
            mix((uint64_t)md.Q); mix((uint64_t)md.D); mix((uint64_t)md.R); mix((uint64_t)md.P); mix((uint64_t)md.parLen);
            { uint64_t pib; memcpy(&pib, &md.pi, 8); mix(pib); }
            mix((uint64_t)ctx->gemm_smem_pad); mix((uint64_t)ctx->force_fail); mix((uint64_t)max_jitter); mix(loop);
            for (auto &sc : subs) {
                mix(sc.cnt); mix((uint64_t)sc.items_max); mix((uint64_t)sc.nstar_max); mix((uint64_t)sc.groups_max);
                for (int t : sc.T) mix((uint64_t)t);
            }
            // Capture on the SECOND sighting (host-buffer entry points only, which synchronise anyway):
            // capturing and instantiating a sequence costs about as much as several direct issues, and
            // callers like the with-update imputation (a new set of history windows at every time stamp)
            // never show the same chunk structure twice.  The device-resident entry points always
            // replay graphs -- their contract is "no host synchronisation inside the call", and
            // without the graph's WHILE node the jitter rounds are driven by the host.
            if (ctx->allow_direct && ctx->graphs.find(key) == ctx->graphs.end()) {
                if (ctx->seen_keys.size() > 8192) ctx->seen_keys.clear();
                if (ctx->seen_keys.insert(key).second) {
                    graph_mode = false;
                    loop = false;
                }
            }
        }
        ctx->retry_on_device = loop;
        if (!graph_mode) {
            const int rc = issue_all();
            if (rc) return rc;
        } else {
            auto it = ctx->graphs.find(key);
            if (it == ctx->graphs.end()) {
                long long before[MEDGP_STAGE_COUNT];
                memcpy(before, ctx->times.launches, sizeof(before));
                cudaGraph_t graph = nullptr;
                GraphEntry ge;
                auto build = [&]() -> int {
                    if (!loop) {  // plain capture of one pass
                        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                        const int rc = issue_all();
                        const cudaError_t ce = cudaStreamEndCapture(st, &graph);  // always: leaves no stream in capture mode
                        if (rc) return rc;
                        if (ce != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return MEDGP_ERR_CUDA; }
                        CU(cudaGraphInstantiate(&ge.exec, graph, 0));
                        return MEDGP_OK;
                    }
                    CU(cudaGraphCreate(&graph, 0));
                    cudaGraphConditionalHandle handle;
                    CU(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
                    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
                    np.type = cudaGraphNodeTypeConditional;
                    np.conditional.handle = handle;
                    np.conditional.type = cudaGraphCondTypeWhile;
                    np.conditional.size = 1;
                    cudaGraphNode_t node;
                    CU(cudaGraphAddNode(&node, graph, nullptr, 0, &np));
                    cudaGraph_t body = np.conditional.phGraph_out[0];
                    CU(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
                    const int rc = issue_all();
                    if (rc == MEDGP_OK)
                        k_retry_decide<<<1, 1024, 0, st>>>(ctx->d_descs + dpos, (int)cnt, ctx->d_fail, max_jitter, handle);
                    cudaGraph_t captured = nullptr;
                    const cudaError_t ce = cudaStreamEndCapture(st, &captured);  // always: leaves no stream in capture mode
                    if (rc) return rc;
                    if (ce != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return MEDGP_ERR_CUDA; }
                    CU(cudaGraphInstantiate(&ge.exec, graph, 0));
                    return MEDGP_OK;
                };
                const int rc = build();
                if (graph) cudaGraphDestroy(graph);
                for (int i = 0; i < MEDGP_STAGE_COUNT; i++) {
                    ge.launches[i] = ctx->times.launches[i] - before[i];
                    ctx->times.launches[i] = before[i];
                }
                if (rc) {
                    cudaGetLastError();
                    return rc;
                }
                if (ctx->graphs.size() >= kGraphCacheMax) {  // bounded cache: evict the least recently used entry
                    auto victim = ctx->graphs.begin();
                    for (auto g = ctx->graphs.begin(); g != ctx->graphs.end(); ++g)
                        if (g->second.last_use < victim->second.last_use) victim = g;
                    cudaStreamSynchronize(st);  // the victim may still be running (evictions are rare)
                    cudaGraphExecDestroy(victim->second.exec);
                    ctx->graphs.erase(victim);
                }
                it = ctx->graphs.emplace(key, ge).first;
            }
            it->second.last_use = ++ctx->graph_clock;
            CU(cudaGraphLaunch(it->second.exec, st));
            for (int i = 0; i < MEDGP_STAGE_COUNT; i++) ctx->times.launches[i] += it->second.launches[i];
        }
        CU(cudaGetLastError());
        dpos += cnt;
    }
    return MEDGP_OK;
}

int check_series_ids(medgp_ctx *ctx, int batch, const int *series_id)
{
    for (int b = 0; b < batch; b++) {
        const int s = series_id[b];
        if (s < 0 || s >= (int)ctx->series.size() || !ctx->series[s].alive) {
            ctx->err = "unknown series id";
            return MEDGP_ERR_ARG;
        }
    }
    return MEDGP_OK;
}

void free_series_mem(Series &s)
{
    s = Series();  // drops its share of the upload's device blob
}

}  // namespace

// ======================================================================================= API

MEDGP_API int medgp_cuda_create(medgp_ctx **out, int device, size_t workspace_bytes)
{
    if (!out) return MEDGP_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
        return MEDGP_ERR_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return MEDGP_ERR_NODEVICE;  // the fatbin holds sm_100a code only
    medgp_ctx *ctx = new medgp_ctx();
    ctx->device = device;
    // The streams that carry a factorisation's critical path (diagonal block -> panel -> first
    // trailing column) run at the highest priority, the auxiliary streams that carry the bulk of
    // the look-ahead trailing updates at the lowest: the block scheduler then hands freed SM slots
    // to a waiting critical kernel before the next wave of a bulk update, also when several
    // matrices are in flight on different streams.  MEDGP_PRIO=0: all streams alike (experiments).
    int prio_lo = 0, prio_hi = 0;
    cudaSetDevice(device);
    if (!(getenv("MEDGP_PRIO") && atoi(getenv("MEDGP_PRIO")) == 0)) cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
        delete ctx;
        return MEDGP_ERR_CUDA;
    }
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (workspace_bytes == 0)
        workspace_bytes = std::min<size_t>((size_t)(0.7 * (double)free_b), (size_t)64 << 30);
    if (cudaMalloc(&ctx->arena, workspace_bytes) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return MEDGP_ERR_NOMEM;
    }
    ctx->arena_bytes = workspace_bytes;
    ctx->max_streams = 8;
    if (const char *ev = getenv("MEDGP_RL")) ctx->force_rl = atoi(ev);
    if (const char *ev = getenv("MEDGP_FOLD_MAX")) ctx->fold_max = atoi(ev);
    if (const char *ev = getenv("MEDGP_STAGGER_US")) ctx->stagger_us = atoi(ev);
    if (const char *ev = getenv("MEDGP_GEMM_SMEM_PAD")) ctx->gemm_smem_pad = atoi(ev);
    if (const char *ev = getenv("MEDGP_CHAIN_DIAG")) ctx->chain_diag = atoi(ev) != 0;
    ctx->timeline = getenv("MEDGP_TIMELINE");
    if (const char *ev = getenv("MEDGP_FUSE_DIAG")) ctx->fuse_diag = atoi(ev) != 0;
    if (const char *ev = getenv("MEDGP_GRAPHS")) ctx->use_graphs = atoi(ev) != 0;
    if (const char *ev = getenv("MEDGP_LAZY_CAPTURE")) ctx->lazy_capture = atoi(ev) != 0;
    if (const char *ev = getenv("MEDGP_DEVICE_RETRY")) ctx->device_retry = atoi(ev) != 0;
    if (const char *ev = getenv("MEDGP_LOOKAHEAD")) ctx->lookahead = atoi(ev) != 0;
    if (const char *ev = getenv("MEDGP_FLOW")) ctx->flow = atoi(ev);
    if (const char *ev = getenv("MEDGP_RL_W")) ctx->rl_width = std::max(0, atoi(ev));
    if (const char *ev = getenv("MEDGP_FORCE_FAIL")) ctx->force_fail = std::max(0, atoi(ev));  // tests of the jitter path through the executables
    if (const char *ev = getenv("MEDGP_STREAMS")) ctx->max_streams = std::max(1, std::min(8, atoi(ev)));
    for (int i = 0; i < 8; i++) {
        cudaStreamCreateWithPriority(&ctx->sub_streams[i], cudaStreamNonBlocking, prio_hi);
        cudaStreamCreateWithPriority(&ctx->aux_streams[i], cudaStreamNonBlocking, prio_lo);
        cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_panel[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_bulk[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    for (int i = 0; i < kDescSlots; i++) cudaEventCreateWithFlags(&ctx->desc_ev[i], cudaEventDisableTiming);
    if (cudaMalloc(&ctx->d_tickets, 8 * kTicketsPerSub * sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_tickets, 0, 8 * kTicketsPerSub * sizeof(int)) != cudaSuccess) {
        medgp_cuda_destroy(ctx);
        return MEDGP_ERR_NOMEM;
    }
    cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    cudaFuncSetAttribute(k_potrf_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    cudaFuncSetAttribute(k_trtri_row, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    cudaFuncSetAttribute(k_lauum, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    cudaFuncSetAttribute(k_potrf_step, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    cudaFuncSetAttribute(k_syrk_update, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    cudaFuncSetAttribute(k_potrf_flow, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    cudaFuncSetAttribute(k_trtri_flow, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    cudaFuncSetAttribute(k_trtri_update, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes + ctx->gemm_smem_pad);
    *out = ctx;
    return MEDGP_OK;
}

MEDGP_API void medgp_cuda_destroy(medgp_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &s : ctx->series)
        if (s.alive) free_series_mem(s);
    resolve_marks(ctx);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second.exec);
    for (int i = 0; i < 8; i++) {
        if (ctx->sub_streams[i]) cudaStreamDestroy(ctx->sub_streams[i]);
        if (ctx->aux_streams[i]) cudaStreamDestroy(ctx->aux_streams[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
        if (ctx->ev_panel[i]) cudaEventDestroy(ctx->ev_panel[i]);
        if (ctx->ev_bulk[i]) cudaEventDestroy(ctx->ev_bulk[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    ctx->blob_pool.clear();
    for (auto &blk : ctx->scg_blocks) {
        cudaFree(blk.dev);
        cudaFreeHost(blk.host);
    }
    if (ctx->h_upload) cudaFreeHost(ctx->h_upload);
    cudaFree(ctx->arena);
    cudaFree(ctx->d_tickets);
    for (int i = 0; i < kDescSlots; i++) {
        if (ctx->h_desc_ring[i]) cudaFreeHost(ctx->h_desc_ring[i]);
        if (ctx->desc_ev[i]) cudaEventDestroy(ctx->desc_ev[i]);
    }
    cudaFree(ctx->d_descs);
    cudaFreeHost(ctx->h_theta); cudaFree(ctx->d_theta);
    cudaFreeHost(ctx->h_out); cudaFree(ctx->d_out);
    cudaFreeHost(ctx->h_status); cudaFree(ctx->d_status); cudaFree(ctx->d_fail);
    cudaFree(ctx->d_star_t); cudaFree(ctx->d_star_meta);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

MEDGP_API const char *medgp_cuda_last_error(const medgp_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : "null context";
}

MEDGP_API int medgp_cuda_model(medgp_ctx *ctx, int Q, int D, int R, double pi_const)
{
    if (!ctx) return MEDGP_ERR_ARG;
    if (Q < 1 || Q > MEDGP_QMAX || D < 1 || R < 1 || Q * D * D > 8192 || !(pi_const > 0)) {
        ctx->err = "model shape out of range (1<=Q<=8, Q*D*D<=8192)";
        return MEDGP_ERR_ARG;
    }
    for (auto &s : ctx->series)
        if (s.alive) {
            ctx->err = "clear the series before changing the model";
            return MEDGP_ERR_ARG;
        }
    cudaSetDevice(ctx->device);
    // captured launch sequences bake in the model (ModelDims by value, Q-templated kernels,
    // shared-memory sizes): none of them survives a change of model
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second.exec);
    ctx->graphs.clear();
    fill_dims(ctx->md, Q, D, R, pi_const);
    const int asm_smem = (Q * D * D + Q) * 8;
    const int npairs = D * (D + 1) / 2;
    const int fin_smem = (Q * D * D + 2 * npairs * Q + D) * 8;
    {
        const int b = std::max(asm_smem, 1024);
        set_assemble_smem_q<1>(b); set_assemble_smem_q<2>(b); set_assemble_smem_q<3>(b); set_assemble_smem_q<4>(b);
        set_assemble_smem_q<5>(b); set_assemble_smem_q<6>(b); set_assemble_smem_q<7>(b); set_assemble_smem_q<8>(b);
    }
    CU(cudaFuncSetAttribute(k_grad_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(fin_smem, 1024)));
    ctx->model_set = true;
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_num_hyp(const medgp_ctx *ctx)
{
    return (ctx && ctx->model_set) ? ctx->md.P : (int)MEDGP_ERR_ARG;
}

namespace {

struct SeriesLayout {  // byte offsets of a series' arrays inside its device blob
    size_t o_t, o_y, o_meta, o_off, o_items, o_pair, o_gs, o_perm, o_fr, total;
};

// Host side of an upload: validate, sort, build the gradient work items / timestamp groups,
// and lay everything out in one blob.
int prepare_series(int D, int n, const int32_t *meta, const float *x, const float *y, int order,
                   Series &s, std::vector<char> &blob, SeriesLayout &lay, std::string &err)
{
    if (n < 1 || (order != MEDGP_ORDER_FEATURE && order != MEDGP_ORDER_TIME && order != MEDGP_ORDER_GIVEN)) {
        err = "add_series: bad argument";
        return MEDGP_ERR_ARG;
    }
    for (int i = 0; i < n; i++)
        if (meta[i] < 0 || meta[i] >= D) {
            err = "add_series: meta out of range";
            return MEDGP_ERR_ARG;
        }
    s.alive = true;
    s.n = n;
    s.npad = (n + MEDGP_NB - 1) / MEDGP_NB * MEDGP_NB;
    s.T = s.npad / MEDGP_NB;
    // feature-major internal order (stable): results are order independent, and the gradient
    // kernel's work items need every feature contiguous.  Time-major (stable) for online
    // imputation: every sliding-window training set is then a leading block plus its group.
    // MEDGP_ORDER_GIVEN keeps the caller's order (the exported factor then refers to it); gradients
    // are available on such a series only when that order happens to be feature-major.
    s.time_order = (order == MEDGP_ORDER_TIME);
    s.given_order = (order == MEDGP_ORDER_GIVEN);
    s.perm.resize(n);
    std::iota(s.perm.begin(), s.perm.end(), 0);
    if (s.time_order)
        std::stable_sort(s.perm.begin(), s.perm.end(), [&](int a, int b) { return x[a] < x[b]; });
    else if (!s.given_order)
        std::stable_sort(s.perm.begin(), s.perm.end(), [&](int a, int b) { return meta[a] < meta[b]; });
    s.grad_ok = !s.time_order;
    if (s.given_order)
        for (int i = 1; i < n; i++)
            if (meta[i] < meta[i - 1]) s.grad_ok = false;
    std::vector<double> ht(s.npad, 0.0), hy(s.npad, 0.0);
    std::vector<int> hm(s.npad, 0), off(D + 1, 0);
    for (int i = 0; i < n; i++) {
        const int src = s.perm[i];
        ht[i] = (double)x[src];
        hy[i] = (double)y[src];
        hm[i] = meta[src];
        off[meta[src] + 1]++;
    }
    for (int d = 0; d < D; d++) off[d + 1] += off[d];
    {
        const auto mm = std::minmax_element(ht.begin(), ht.begin() + n);
        s.trange2 = (*mm.second - *mm.first) * (*mm.second - *mm.first);
    }
    // gradient work items: (block of 32 rows) x (column range [jb, je) cut at feature boundaries,
    // about kGradCols columns, never past the block's last row); a row block's rows split into
    // segments of equal feature, numbered globally in row order, so the segments of feature d
    // are seg_start[d] .. seg_start[d+1]-1.
    std::vector<int4> items;
    std::vector<int> seg_start(D + 1, 0);
    int nseg = 0;
    for (int i0 = 0; i0 < n && s.grad_ok; i0 += kGradRows) {
        const int i1 = std::min(i0 + kGradRows, n);
        int jb = 0;
        for (int f = 0; f < D && off[f] < i1; f++) {
            const int je = std::min(off[f + 1], i1);
            if (je - jb >= kGradCols || je == i1) {
                if (je > jb) items.push_back(make_int4(i0 / kGradRows, jb, je, nseg));
                jb = je;
            }
        }
        for (int i = i0; i < i1; i++)
            if (i == i0 || hm[i] != hm[i - 1]) {
                seg_start[hm[i] + 1]++;
                nseg++;
            }
    }
    for (int d = 0; d < D; d++) seg_start[d + 1] += seg_start[d];
    s.nitems = (int)items.size();
    s.nseg = nseg;
    // same-timestamp groups (float equality, as main_one_test.cpp:300 compares)
    std::vector<int> gstart;
    if (s.time_order) {
        for (int i = 0; i < n; i++)
            if (i == 0 || x[s.perm[i]] != x[s.perm[i - 1]]) gstart.push_back(i);
        gstart.push_back(n);
        s.ngroups = (int)gstart.size() - 1;
        for (int g = 0; g < s.ngroups; g++)
            if (gstart[g + 1] - gstart[g] > MEDGP_GMAX) {
                err = "add_series: more than 32 points share one timestamp (online imputation groups)";
                return MEDGP_ERR_ARG;
            }
    }
    lay.o_t = 0;
    lay.o_y = lay.o_t + (size_t)s.npad * 8;
    lay.o_meta = lay.o_y + (size_t)s.npad * 8;
    lay.o_off = align_up(lay.o_meta + (size_t)s.npad * 4, 16);
    lay.o_items = align_up(lay.o_off + (size_t)(D + 1) * 4, 16);
    lay.o_pair = lay.o_items + std::max<size_t>(1, items.size()) * sizeof(int4);
    lay.o_gs = align_up(lay.o_pair + seg_start.size() * sizeof(int), 16);
    lay.o_perm = lay.o_gs + gstart.size() * sizeof(int);
    lay.o_fr = align_up(lay.o_perm + (s.time_order ? (size_t)n * sizeof(int) : 0), 16);
    lay.total = lay.o_fr + (size_t)s.T * sizeof(int2);
    const size_t base = blob.size();  // 256-aligned by the callers
    blob.resize(base + align_up(lay.total, 256), 0);
    char *bp = blob.data() + base;
    memcpy(bp + lay.o_t, ht.data(), (size_t)s.npad * 8);
    memcpy(bp + lay.o_y, hy.data(), (size_t)s.npad * 8);
    memcpy(bp + lay.o_meta, hm.data(), (size_t)s.npad * 4);
    memcpy(bp + lay.o_off, off.data(), (size_t)(D + 1) * 4);
    if (!items.empty()) memcpy(bp + lay.o_items, items.data(), items.size() * sizeof(int4));
    memcpy(bp + lay.o_pair, seg_start.data(), seg_start.size() * sizeof(int));
    if (s.time_order) {
        memcpy(bp + lay.o_gs, gstart.data(), gstart.size() * sizeof(int));
        memcpy(bp + lay.o_perm, s.perm.data(), (size_t)n * sizeof(int));
    }
    // feature range of every 64-point block: the assembly kernel stages only that part of B_q
    std::vector<int2> fr(s.T);
    for (int b = 0; b < s.T; b++) {
        int lo = D - 1, hi = 0;
        for (int i = b * MEDGP_NB; i < std::min(n, (b + 1) * MEDGP_NB); i++) {
            lo = std::min(lo, hm[i]);
            hi = std::max(hi, hm[i]);
        }
        fr[b] = make_int2(std::min(lo, hi), hi);
    }
    memcpy(bp + lay.o_fr, fr.data(), fr.size() * sizeof(int2));
    return MEDGP_OK;
}

void bind_series(Series &s, const SeriesLayout &lay, char *d_base, const std::shared_ptr<DeviceBlob> &blob)
{
    s.blob = blob;
    s.d_t = (double *)(d_base + lay.o_t);
    s.d_y = (double *)(d_base + lay.o_y);
    s.d_meta = (int *)(d_base + lay.o_meta);
    s.d_off = (int *)(d_base + lay.o_off);
    s.d_items = (int4 *)(d_base + lay.o_items);
    s.d_seg_start = (int *)(d_base + lay.o_pair);
    s.d_frange = (int2 *)(d_base + lay.o_fr);
    if (s.time_order) {
        s.d_gstart = (int *)(d_base + lay.o_gs);
        s.d_perm = (int *)(d_base + lay.o_perm);
    }
}

int place_series(medgp_ctx *ctx, Series &&s)
{
    int id;
    if (!ctx->free_slots.empty()) {  // reuse a dead slot
        id = ctx->free_slots.back();
        ctx->free_slots.pop_back();
    } else {
        id = (int)ctx->series.size();
        ctx->series.emplace_back();
    }
    ctx->series[id] = std::move(s);
    return id;
}

}  // namespace

// One device allocation and ONE host-to-device copy for `count` series (test-time workloads
// upload a training window per patient and time stamp: thousands of short series).
MEDGP_API int medgp_cuda_add_series_batch(medgp_ctx *ctx, int count, const int *n, const int32_t *meta, const float *x,
                                          const float *y, int order, int *out_series_ids)
{
    if (!ctx || !ctx->model_set || count < 0 || !n || !meta || !x || !y || !out_series_ids) {
        if (ctx) ctx->err = "add_series_batch: bad argument or model not set";
        return MEDGP_ERR_ARG;
    }
    if (count == 0) return MEDGP_OK;
    cudaSetDevice(ctx->device);
    std::vector<Series> ser(count);
    std::vector<SeriesLayout> lay(count);
    std::vector<size_t> base(count), first(count);
    std::vector<std::vector<char> > part(count);
    size_t pos = 0;
    for (int b = 0; b < count; b++) {
        if (n[b] < 1) { ctx->err = "add_series_batch: empty series"; return MEDGP_ERR_ARG; }
        first[b] = pos;
        pos += (size_t)n[b];
    }
    // sorting and the work-item lists of the series are independent: spread them over the host cores
    int bad = MEDGP_OK;
#pragma omp parallel for schedule(dynamic, 16) if (count >= 64)
    for (int b = 0; b < count; b++) {
        std::string err;
        const int rc = prepare_series(ctx->md.D, n[b], meta + first[b], x + first[b], y + first[b], order, ser[b], part[b], lay[b], err);
        if (rc) {
#pragma omp critical
            { bad = rc; ctx->err = err; }
        }
    }
    if (bad) return bad;
    size_t total = 0;
    for (int b = 0; b < count; b++) { base[b] = total; total += part[b].size(); }
    if (total > ctx->upload_cap) {
        if (ctx->h_upload) cudaFreeHost(ctx->h_upload);
        ctx->h_upload = nullptr;
        ctx->upload_cap = 0;
        CU(cudaMallocHost(&ctx->h_upload, total + total / 2));
        ctx->upload_cap = total + total / 2;
    }
    char *blob = ctx->h_upload;
#pragma omp parallel for schedule(static) if (count >= 64)
    for (int b = 0; b < count; b++) memcpy(blob + base[b], part[b].data(), part[b].size());
    size_t cap = 0;
    char *d_blob = ctx->blob_pool.take(total, cap);
    if (!d_blob) { ctx->err = "add_series: out of device memory"; return MEDGP_ERR_NOMEM; }
    std::shared_ptr<DeviceBlob> owner = std::make_shared<DeviceBlob>(d_blob, cap, &ctx->blob_pool);
    // on the context's stream: ordered behind whatever still reads a recycled blob
    CU(cudaMemcpyAsync(d_blob, blob, total, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < count; b++) {
        bind_series(ser[b], lay[b], d_blob + base[b], owner);
        out_series_ids[b] = place_series(ctx, std::move(ser[b]));
    }
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_add_series_ordered(medgp_ctx *ctx, int n, const int32_t *meta, const float *x,
                                            const float *y, int order, int *out_series_id)
{
    if (!ctx || !ctx->model_set || n < 1 || !meta || !x || !y || !out_series_id) {
        if (ctx) ctx->err = "add_series: bad argument or model not set";
        return MEDGP_ERR_ARG;
    }
    return medgp_cuda_add_series_batch(ctx, 1, &n, meta, x, y, order, out_series_id);
}

MEDGP_API int medgp_cuda_add_series(medgp_ctx *ctx, int n, const int32_t *meta, const float *x,
                                    const float *y, int *out_series_id)
{
    return medgp_cuda_add_series_ordered(ctx, n, meta, x, y, MEDGP_ORDER_FEATURE, out_series_id);
}

MEDGP_API int medgp_cuda_free_series(medgp_ctx *ctx, int series_id)
{
    if (!ctx || series_id < 0 || series_id >= (int)ctx->series.size() || !ctx->series[series_id].alive)
        return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_series_mem(ctx->series[series_id]);
    ctx->free_slots.push_back(series_id);
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_free_series_batch(medgp_ctx *ctx, int count, const int *series_ids)
{
    if (!ctx || count < 0 || (count > 0 && !series_ids)) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    int rc = check_series_ids(ctx, count, series_ids);
    if (rc) return rc;
    cudaStreamSynchronize(ctx->stream);
    for (int b = 0; b < count; b++) {
        if (!ctx->series[series_ids[b]].alive) continue;  // listed twice
        free_series_mem(ctx->series[series_ids[b]]);
        ctx->free_slots.push_back(series_ids[b]);
    }
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_clear_series(medgp_ctx *ctx)
{
    if (!ctx) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &s : ctx->series)
        if (s.alive) free_series_mem(s);
    ctx->series.clear();
    ctx->free_slots.clear();
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_sync(medgp_ctx *ctx)
{
    if (!ctx) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    resolve_marks(ctx);
    return MEDGP_OK;
}

// host-driven jitter rounds for the call paths that run without the graph's WHILE node
// (profiling, MEDGP_GRAPHS=0, MEDGP_DEVICE_RETRY=0): re-run the failed evaluations with one more
// noise addition until they pass or kMaxJitter is reached (c_inference_exact.cpp:99-108)
static int host_retry_rounds(medgp_ctx *ctx, std::vector<Request> reqs, int batch, const double *d_theta, int mode,
                             double *d_nlml, double *d_grad, int *d_status, double *d_mean, double *d_var)
{
    cudaStream_t st = ctx->stream;
    for (int round = 1; round <= kMaxJitter; round++) {
        CU(cudaMemcpyAsync(ctx->h_status, d_status, batch * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        resolve_marks(ctx);
        std::vector<Request> again;
        for (auto &rq : reqs)
            if (ctx->h_status[rq.out_index] < 0 && rq.jitter < kMaxJitter) {
                Request r2 = rq;
                r2.jitter++;
                again.push_back(r2);
            }
        if (again.empty()) break;
        reqs.swap(again);
        CU(cudaMemsetAsync(ctx->d_fail, 0, batch * sizeof(int), st));
        const int rc = run_batch(ctx, reqs, d_theta, mode, d_nlml, d_grad, d_status, d_mean, d_var, 0);
        if (rc) return rc;
    }
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_nlml_grad_device(medgp_ctx *ctx, int batch, const int *series_id,
                                          const double *d_theta, int want_grad, double *d_nlml,
                                          double *d_grad, int *d_status)
{
    if (!ctx || !ctx->model_set || batch < 0 || !series_id || !d_theta || !d_nlml || !d_status ||
        (want_grad && !d_grad)) {
        if (ctx) ctx->err = "nlml_grad_device: bad argument";
        return MEDGP_ERR_ARG;
    }
    if (batch == 0) return MEDGP_OK;
    cudaSetDevice(ctx->device);
    int rc = check_series_ids(ctx, batch, series_id);
    if (rc) return rc;
    if (want_grad)
        for (int b = 0; b < batch; b++)
            if (!ctx->series[series_id[b]].grad_ok) {
                ctx->err = "nlml_grad_device: gradients need a feature-ordered series (medgp_cuda_add_series)";
                return MEDGP_ERR_ARG;
            }
    if (ctx->profile) {  // stage marks are resolved per call
        CU(cudaStreamSynchronize(ctx->stream));
        resolve_marks(ctx);
    }
    // no host synchronisation here: the call fills its own descriptor slot, and everything it
    // enqueues is ordered behind the previous call on the context's stream
    rc = ensure_staging(ctx, batch, 0);
    if (rc) return rc;
    std::vector<Request> reqs(batch);
    for (int b = 0; b < batch; b++) reqs[b] = {series_id[b], b, 0, 0, 0};
    CU(cudaMemsetAsync(ctx->d_fail, 0, batch * sizeof(int), ctx->stream));
    rc = run_batch(ctx, reqs, d_theta, want_grad ? 1 : 0, d_nlml, d_grad, d_status, nullptr, nullptr, 0);
    if (rc) return rc;
    if (!ctx->retry_on_device) {
        rc = host_retry_rounds(ctx, std::move(reqs), batch, d_theta, want_grad ? 1 : 0, d_nlml, d_grad, d_status, nullptr, nullptr);
        if (rc) return rc;
    }
    return release_desc_slot(ctx);
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

MEDGP_API int medgp_cuda_host_alloc(medgp_ctx *ctx, size_t bytes, void **h_ptr)
{
    if (!ctx || !h_ptr) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_host_free(medgp_ctx *ctx, void *h_ptr)
{
    if (!ctx) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaFreeHost(h_ptr));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_nlml_grad(medgp_ctx *ctx, int batch, const int *series_id,
                                   const double *theta, int want_grad, double *nlml, double *grad,
                                   int *status)
{
    if (!ctx || !ctx->model_set || batch < 0 || !series_id || !theta || !nlml || !status ||
        (want_grad && !grad)) {
        if (ctx) ctx->err = "nlml_grad: bad argument";
        return MEDGP_ERR_ARG;
    }
    if (batch == 0) return MEDGP_OK;
    cudaSetDevice(ctx->device);
    int rc = check_series_ids(ctx, batch, series_id);
    if (rc) return rc;
    if (want_grad)
        for (int b = 0; b < batch; b++)
            if (!ctx->series[series_id[b]].grad_ok) {
                ctx->err = "nlml_grad: gradients need a feature-ordered series (medgp_cuda_add_series)";
                return MEDGP_ERR_ARG;
            }
    CU(cudaStreamSynchronize(ctx->stream));
    rc = ensure_staging(ctx, batch, 0);
    if (rc) return rc;
    const size_t P = ctx->md.P;
    cudaStream_t st = ctx->stream;
    // page-locked caller buffers are copied to/from directly; pageable ones go through the
    // context's pinned staging buffers
    const bool pin_in = is_pinned_host(theta);
    const bool pin_out = is_pinned_host(nlml) && is_pinned_host(status) && (!want_grad || is_pinned_host(grad));
    if (!pin_in) memcpy(ctx->h_theta, theta, (size_t)batch * P * 8);
    CU(cudaMemcpyAsync(ctx->d_theta, pin_in ? theta : ctx->h_theta, (size_t)batch * P * 8, cudaMemcpyHostToDevice, st));
    double *d_nlml = ctx->d_out, *d_grad = ctx->d_out + batch;
    auto fetch_outputs = [&]() -> int {
        if (pin_out) {
            CU(cudaMemcpyAsync(nlml, d_nlml, (size_t)batch * 8, cudaMemcpyDeviceToHost, st));
            if (want_grad) CU(cudaMemcpyAsync(grad, d_grad, (size_t)batch * P * 8, cudaMemcpyDeviceToHost, st));
        } else {
            const size_t nout = want_grad ? (size_t)batch * (P + 1) : (size_t)batch;
            CU(cudaMemcpyAsync(ctx->h_out, ctx->d_out, nout * 8, cudaMemcpyDeviceToHost, st));
        }
        return MEDGP_OK;
    };
    std::vector<Request> reqs(batch);
    for (int b = 0; b < batch; b++) reqs[b] = {series_id[b], b, 0, 0, 0};
    CU(cudaMemsetAsync(ctx->d_fail, 0, batch * sizeof(int), st));
    ctx->allow_direct = ctx->lazy_capture;  // a chunk structure seen for the first time is issued directly
    rc = run_batch(ctx, reqs, ctx->d_theta, want_grad ? 1 : 0, d_nlml, d_grad, ctx->d_status, nullptr, nullptr, 0);
    ctx->allow_direct = false;
    if (rc) return rc;
    // jitter retries happened inside the launch sequence (graph WHILE node); without it the
    // host drives them
    if (!ctx->retry_on_device) {
        rc = host_retry_rounds(ctx, std::move(reqs), batch, ctx->d_theta, want_grad ? 1 : 0, d_nlml, d_grad,
                               ctx->d_status, nullptr, nullptr);
        if (rc) return rc;
    }
    rc = release_desc_slot(ctx);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->h_status, ctx->d_status, batch * sizeof(int), cudaMemcpyDeviceToHost, st));
    rc = fetch_outputs();
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));  // the one wait of the call
    resolve_marks(ctx);
    if (!pin_out) {
        memcpy(nlml, ctx->h_out, (size_t)batch * 8);
        if (want_grad) memcpy(grad, ctx->h_out + batch, (size_t)batch * P * 8);
    }
    memcpy(status, ctx->h_status, (size_t)batch * sizeof(int));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_predict(medgp_ctx *ctx, int batch, const int *series_id,
                                 const double *theta, const int *star_offset,
                                 const int32_t *meta_star, const float *x_star, double *mean,
                                 double *var, int *status)
{
    if (!ctx || !ctx->model_set || batch < 0 || !series_id || !theta || !star_offset ||
        !meta_star || !x_star || !mean || !var || !status) {
        if (ctx) ctx->err = "predict: bad argument";
        return MEDGP_ERR_ARG;
    }
    if (batch == 0) return MEDGP_OK;
    cudaSetDevice(ctx->device);
    int rc = check_series_ids(ctx, batch, series_id);
    if (rc) return rc;
    const int nstar = star_offset[batch];
    for (int b = 0; b < batch; b++)
        if (star_offset[b + 1] < star_offset[b]) { ctx->err = "predict: star_offset not monotone"; return MEDGP_ERR_ARG; }
    for (int i = 0; i < nstar; i++)
        if (meta_star[i] < 0 || meta_star[i] >= ctx->md.D) { ctx->err = "predict: meta_star out of range"; return MEDGP_ERR_ARG; }
    CU(cudaStreamSynchronize(ctx->stream));
    rc = ensure_staging(ctx, batch, nstar);
    if (rc) return rc;
    const size_t P = ctx->md.P;
    cudaStream_t st = ctx->stream;
    memcpy(ctx->h_theta, theta, (size_t)batch * P * 8);
    CU(cudaMemcpyAsync(ctx->d_theta, ctx->h_theta, (size_t)batch * P * 8, cudaMemcpyHostToDevice, st));
    if (nstar > 0) {
        std::vector<double> ts(nstar);
        for (int i = 0; i < nstar; i++) ts[i] = (double)x_star[i];
        CU(cudaMemcpyAsync(ctx->d_star_t, ts.data(), nstar * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_star_meta, meta_star, nstar * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));  // ts is a pageable temporary
    }
    double *d_nlml = ctx->d_out, *d_mean = ctx->d_out + batch, *d_var = d_mean + nstar;
    std::vector<Request> reqs(batch);
    for (int b = 0; b < batch; b++)
        reqs[b] = {series_id[b], b, 0, star_offset[b + 1] - star_offset[b], star_offset[b]};
    CU(cudaMemsetAsync(ctx->d_fail, 0, batch * sizeof(int), st));
    ctx->allow_direct = ctx->lazy_capture;  // a chunk structure seen for the first time is issued directly
    rc = run_batch(ctx, reqs, ctx->d_theta, 2, d_nlml, nullptr, ctx->d_status, d_mean, d_var, 0);
    ctx->allow_direct = false;
    if (rc) return rc;
    if (!ctx->retry_on_device) {
        rc = host_retry_rounds(ctx, std::move(reqs), batch, ctx->d_theta, 2, d_nlml, nullptr, ctx->d_status, d_mean, d_var);
        if (rc) return rc;
    }
    rc = release_desc_slot(ctx);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->h_status, ctx->d_status, batch * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_out, ctx->d_out, ((size_t)batch + 2 * (size_t)nstar) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    resolve_marks(ctx);
    memcpy(mean, ctx->h_out + batch, (size_t)nstar * 8);
    memcpy(var, ctx->h_out + batch + nstar, (size_t)nstar * 8);
    memcpy(status, ctx->h_status, (size_t)batch * sizeof(int));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_predict_online(medgp_ctx *ctx, int batch, const int *series_id,
                                        const double *theta, double *mean, double *var, int *status)
{
    if (!ctx || !ctx->model_set || batch < 0 || !series_id || !theta || !mean || !var || !status) {
        if (ctx) ctx->err = "predict_online: bad argument";
        return MEDGP_ERR_ARG;
    }
    if (batch == 0) return MEDGP_OK;
    cudaSetDevice(ctx->device);
    int rc = check_series_ids(ctx, batch, series_id);
    if (rc) return rc;
    size_t ntot = 0;
    std::vector<Request> reqs(batch);
    for (int b = 0; b < batch; b++) {
        const Series &s = ctx->series[series_id[b]];
        if (!s.time_order) {
            ctx->err = "predict_online: series must be uploaded with MEDGP_ORDER_TIME";
            return MEDGP_ERR_ARG;
        }
        reqs[b] = {series_id[b], b, 0, 0, (int)ntot};
        ntot += (size_t)s.n;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    rc = ensure_staging(ctx, batch, ntot);
    if (rc) return rc;
    const size_t P = ctx->md.P;
    cudaStream_t st = ctx->stream;
    memcpy(ctx->h_theta, theta, (size_t)batch * P * 8);
    CU(cudaMemcpyAsync(ctx->d_theta, ctx->h_theta, (size_t)batch * P * 8, cudaMemcpyHostToDevice, st));
    double *d_nlml = ctx->d_out, *d_mean = ctx->d_out + batch, *d_var = d_mean + ntot;
    CU(cudaMemsetAsync(ctx->d_fail, 0, batch * sizeof(int), st));
    ctx->allow_direct = ctx->lazy_capture;  // a chunk structure seen for the first time is issued directly
    rc = run_batch(ctx, reqs, ctx->d_theta, 3, d_nlml, nullptr, ctx->d_status, d_mean, d_var, 0);
    ctx->allow_direct = false;
    if (rc) return rc;
    rc = release_desc_slot(ctx);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->h_status, ctx->d_status, batch * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_out, ctx->d_out, ((size_t)batch + 2 * ntot) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    resolve_marks(ctx);
    memcpy(mean, ctx->h_out + batch, ntot * 8);
    memcpy(var, ctx->h_out + batch + ntot, ntot * 8);
    memcpy(status, ctx->h_status, (size_t)batch * sizeof(int));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_debug_matrices(medgp_ctx *ctx, int series_id, const double *theta,
                                        double *K, double *L, double *alpha, double *Kinv)
{
    if (!ctx || !ctx->model_set || !theta) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    int rc = check_series_ids(ctx, 1, &series_id);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    rc = ensure_staging(ctx, 1, 0);
    if (rc) return rc;
    const Series &s = ctx->series[series_id];
    if (!s.grad_ok && (alpha || Kinv)) {
        ctx->err = "debug_matrices: alpha / K^-1 need a feature-ordered series (they come from the gradient path)";
        return MEDGP_ERR_ARG;
    }
    const ModelDims &md = ctx->md;
    cudaStream_t st = ctx->stream;
    const size_t np = s.npad, n = s.n;
    CU(cudaMemcpyAsync(ctx->d_theta, theta, (size_t)md.P * 8, cudaMemcpyHostToDevice, st));
    std::vector<double> hM(np * np), ha(np), hT((size_t)s.T * s.T * kTileElems);
    auto fetch = [&]() -> int {
        // the single evaluation's M is the first arena allocation; convert tile-major ->
        // plain column-major (ld = np)
        CU(cudaMemcpyAsync(hT.data(), ctx->arena, hT.size() * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (size_t j = 0; j < np; j++)
            for (size_t i = 0; i < np; i++) hM[j * np + i] = hT[elem_off(s.T, (int)i, (int)j)];
        return MEDGP_OK;
    };
    std::vector<Request> one = {{series_id, 0, 0, 0, 0}};
    // the taps read buffers between stages and re-run single kernels on the descriptors: no
    // device-side jitter loop here (it would retire the descriptor), one pass only
    struct RetryOff {
        medgp_ctx *c; bool saved;
        explicit RetryOff(medgp_ctx *cc) : c(cc), saved(cc->device_retry) { c->device_retry = false; }
        ~RetryOff() { c->device_retry = saved; }
    } retry_off(ctx);
    if (K) {
        // a full NLML pass sets up descriptors and parameters ...
        rc = run_batch(ctx, one, ctx->d_theta, 0, ctx->d_out, nullptr, ctx->d_status, nullptr, nullptr, 0);
        if (rc) return rc;
        // ... then assembly alone is re-run into the same buffer (potrf overwrote it)
        const int asm_smem = (md.Q * md.D * md.D + md.Q) * 8;
        launch_assemble(md.Q, dim3(s.T * (s.T + 1) / 2, 1), asm_smem, st, ctx->d_descs, md);
        if ((rc = fetch())) return rc;
        for (size_t i = 0; i < n; i++)
            for (size_t j = 0; j <= i; j++) {
                const double v = hM[j * np + i];
                K[(size_t)s.perm[i] * n + s.perm[j]] = v;
                K[(size_t)s.perm[j] * n + s.perm[i]] = v;
            }
    }
    if (L) {
        CU(cudaMemsetAsync(ctx->d_fail, 0, sizeof(int), st));
        rc = run_batch(ctx, one, ctx->d_theta, 0, ctx->d_out, nullptr, ctx->d_status, nullptr, nullptr, 0);
        if (rc) return rc;
        if ((rc = fetch())) return rc;
        for (size_t i = 0; i < n; i++)
            for (size_t j = 0; j < n; j++) L[i * n + j] = (j <= i) ? hM[j * np + i] : 0.0;
    }
    if (alpha || Kinv) {
        CU(cudaMemsetAsync(ctx->d_fail, 0, sizeof(int), st));
        rc = run_batch(ctx, one, ctx->d_theta, 1, ctx->d_out, ctx->d_out + 1, ctx->d_status, nullptr, nullptr, 0);
        if (rc) return rc;
        if ((rc = fetch())) return rc;
        if (Kinv)
            for (size_t i = 0; i < n; i++)
                for (size_t j = 0; j <= i; j++) {
                    const double v = hM[j * np + i];
                    Kinv[(size_t)s.perm[i] * n + s.perm[j]] = v;
                    Kinv[(size_t)s.perm[j] * n + s.perm[i]] = v;
                }
        if (alpha) {
            CU(cudaMemcpy(ha.data(), ctx->h_descs[0].alpha, np * 8, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; i++) alpha[s.perm[i]] = ha[i];
        }
    }
    return MEDGP_OK;
}


// ======================================================================= optimiser sessions
struct medgp_scg {
    medgp_ctx *ctx = nullptr;
    ScgSession S{};
    std::vector<int> series;
    signed char *d_ptype = nullptr, *d_pexp = nullptr;
    float *d_ppar = nullptr;
    int *h_active = nullptr;  // pinned: [0] = live count, [1 + b] = skip flag of instance b
    double *d_theta0 = nullptr;  // staging of scg_start
    int *d_len = nullptr;
    int block = -1;              // index of the session's block in ctx->scg_blocks
    std::vector<int> live;    // instances that wanted an evaluation at the last poll
    std::vector<int> launch;  // instances the super-steps are launched for (a superset of live)
    bool started = false;
};

MEDGP_API int medgp_cuda_scg_create(medgp_ctx *ctx, int count, medgp_scg **out)
{
    if (!ctx || !ctx->model_set || count < 1 || !out) {
        if (ctx) ctx->err = "scg_create: bad argument or model not set";
        return MEDGP_ERR_ARG;
    }
    cudaSetDevice(ctx->device);
    const size_t P = ctx->md.P, n = (size_t)count;
    // one device block per session, carved below; the last two pieces stage theta0 / budgets of scg_start
    const size_t sizes[] = {n * sizeof(ScgScalars), n * SCG_NVEC * P * 8, n * P * 8, n * 8, n * P * 8, n * sizeof(int),
                            n * sizeof(int), sizeof(int), n * P, n * P, n * P * 2 * sizeof(float), n * P * 8, n * sizeof(int)};
    size_t total = 0;
    for (size_t b : sizes) total += align_up(b, 256);
    const size_t host_bytes = (n + 1) * sizeof(int);
    int slot = -1;
    for (size_t k = 0; k < ctx->scg_blocks.size(); k++) {
        const auto &blk = ctx->scg_blocks[k];
        if (!blk.busy && blk.dev_bytes >= total && blk.host_bytes >= host_bytes &&
            (slot < 0 || blk.dev_bytes < ctx->scg_blocks[slot].dev_bytes))
            slot = (int)k;
    }
    if (slot < 0) {
        medgp_ctx::ScgBlock blk;
        if (cudaMalloc(&blk.dev, total) != cudaSuccess || cudaMallocHost(&blk.host, host_bytes) != cudaSuccess) {
            if (blk.dev) cudaFree(blk.dev);
            cudaGetLastError();
            ctx->err = "scg_create: out of memory";
            return MEDGP_ERR_NOMEM;
        }
        blk.dev_bytes = total;
        blk.host_bytes = host_bytes;
        ctx->scg_blocks.push_back(blk);
        slot = (int)ctx->scg_blocks.size() - 1;
    }
    ctx->scg_blocks[slot].busy = true;
    medgp_scg *g = new medgp_scg();
    g->ctx = ctx;
    g->block = slot;
    g->S.count = count;
    g->S.P = (int)P;
    char *p = ctx->scg_blocks[slot].dev;
    int piece = 0;
    auto take = [&]() {
        char *r = p;
        p += align_up(sizes[piece++], 256);
        return r;
    };
    g->S.sc = (ScgScalars *)take();
    g->S.vec = (double *)take();
    g->S.theta = (double *)take();
    g->S.nlml = (double *)take();
    g->S.grad = (double *)take();
    g->S.status = (int *)take();
    g->S.skip = (int *)take();
    g->S.active = (int *)take();
    g->d_ptype = (signed char *)take();
    g->d_pexp = (signed char *)take();
    g->d_ppar = (float *)take();
    g->d_theta0 = (double *)take();
    g->d_len = (int *)take();
    g->h_active = ctx->scg_blocks[slot].host;
    *out = g;
    return MEDGP_OK;
}

MEDGP_API void medgp_cuda_scg_destroy(medgp_scg *g)
{
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);  // nothing enqueued for the session is still running
    if (g->block >= 0 && g->block < (int)g->ctx->scg_blocks.size()) g->ctx->scg_blocks[g->block].busy = false;
    delete g;
}

static int scg_count_active(medgp_scg *g, int *active_left);

MEDGP_API int medgp_cuda_scg_start(medgp_scg *g, const int *series_id, const double *theta0, const int *max_iteration,
                                   const signed char *prior_type, const signed char *prior_exp, const float *prior_param)
{
    if (!g || !series_id || !theta0 || !max_iteration || ((prior_type != nullptr) != (prior_param != nullptr)) ||
        ((prior_type != nullptr) != (prior_exp != nullptr))) {
        if (g) g->ctx->err = "scg_start: bad argument";
        return MEDGP_ERR_ARG;
    }
    medgp_ctx *ctx = g->ctx;
    cudaSetDevice(ctx->device);
    const int count = g->S.count;
    const size_t P = ctx->md.P;
    int rc = check_series_ids(ctx, count, series_id);
    if (rc) return rc;
    for (int b = 0; b < count; b++) {
        if (!ctx->series[series_id[b]].grad_ok) {
            ctx->err = "scg_start: gradients need a feature-ordered series (medgp_cuda_add_series)";
            return MEDGP_ERR_ARG;
        }
        if (prior_type)
            for (size_t k = 0; k < P; k++) {
                const int t = prior_type[(size_t)b * P + k];
                if (t < -1 || t > 2) {
                    ctx->err = "scg_start: prior types -1 (none), 0 (clamp), 1 (normal), 2 (laplace) run on the device";
                    return MEDGP_ERR_ARG;
                }
            }
    }
    cudaStream_t st = ctx->stream;
    CU(cudaStreamSynchronize(st));  // the staging buffers below are plain temporaries
    g->series.assign(series_id, series_id + count);
    double *d_theta0 = g->d_theta0;
    int *d_len = g->d_len;
    CU(cudaMemcpyAsync(d_theta0, theta0, (size_t)count * P * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_len, max_iteration, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, st));
    if (prior_type) {
        CU(cudaMemcpyAsync(g->d_ptype, prior_type, (size_t)count * P, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(g->d_pexp, prior_exp, (size_t)count * P, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(g->d_ppar, prior_param, (size_t)count * P * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
        g->S.ptype = g->d_ptype; g->S.pexp = g->d_pexp; g->S.ppar = g->d_ppar;
    } else {
        g->S.ptype = nullptr; g->S.pexp = nullptr; g->S.ppar = nullptr;
    }
    k_scg_start<<<count, MEDGP_SCG_THREADS, 0, st>>>(g->S, d_theta0, d_len);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));  // the caller's host arrays have been consumed
    g->started = true;
    g->launch.clear();
    return scg_count_active(g, nullptr);
}

// the tail of a super-step (and of the test tap): advance every live instance
static int scg_advance(medgp_scg *g)
{
    medgp_ctx *ctx = g->ctx;
    k_scg_advance<<<g->S.count, MEDGP_SCG_THREADS, 0, ctx->stream>>>(g->S, ctx->md);
    CU(cudaGetLastError());
    return MEDGP_OK;
}

static int scg_count_active(medgp_scg *g, int *active_left)
{
    medgp_ctx *ctx = g->ctx;
    k_scg_count<<<1, 256, 0, ctx->stream>>>(g->S);
    CU(cudaMemcpyAsync(g->h_active, g->S.active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(g->h_active + 1, g->S.skip, (size_t)g->S.count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    resolve_marks(ctx);
    // the instances still running: the next super-steps are launched for these only (finished
    // ones are still passed over on the device if they finish between two polls)
    g->live.clear();
    for (int b = 0; b < g->S.count; b++)
        if (!g->h_active[1 + b]) g->live.push_back(b);
    // Finished instances are passed over on the device, so the launch set only has to CONTAIN the
    // live ones.  It is narrowed when enough of it has finished (a new set means a new captured
    // launch sequence, so not at every poll): MEDGP_SCG_COMPACT = fraction of the launch set that
    // must still be live to keep it (default 0.6; 0 never narrows, 1 narrows at every poll).
    static const double keep = getenv("MEDGP_SCG_COMPACT") ? atof(getenv("MEDGP_SCG_COMPACT")) : 0.6;
    if (g->launch.empty() || (double)g->live.size() < keep * (double)g->launch.size() || g->live.empty()) g->launch = g->live;
    if (active_left) *active_left = *g->h_active;
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_scg_run(medgp_scg *g, int super_steps, int *active_left)
{
    if (!g || !g->started || super_steps < 0) {
        if (g) g->ctx->err = "scg_run: session not started";
        return MEDGP_ERR_ARG;
    }
    medgp_ctx *ctx = g->ctx;
    cudaSetDevice(ctx->device);
    const int count = g->S.count;
    std::vector<Request> reqs;
    for (int b : g->launch) reqs.push_back({g->series[b], b, 0, 0, 0});
    // MEDGP_SCG_TRACE: one line per poll on stderr (launch set, live instances, device time of the
    // super-steps, host time to enqueue them, wall time of the whole poll)
    static const bool trace = getenv("MEDGP_SCG_TRACE") != nullptr;
    static cudaEvent_t tr_a = nullptr, tr_b = nullptr;
    const auto w0 = std::chrono::steady_clock::now();
    if (trace) {
        if (!tr_a) { cudaEventCreate(&tr_a); cudaEventCreate(&tr_b); }
        cudaEventRecord(tr_a, ctx->stream);
    }
    struct TraceEnd {
        medgp_scg *g; bool on; std::chrono::steady_clock::time_point w0; size_t nreq; int steps;
        std::chrono::steady_clock::time_point w1;
        ~TraceEnd() {
            if (!on) return;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, tr_a, tr_b);
            const auto w2 = std::chrono::steady_clock::now();
            fprintf(stderr, "scg poll: launch=%zu live=%zu steps=%d gpu=%.3f ms enqueue=%.3f ms wall=%.3f ms\n", nreq,
                    g->live.size(), steps, ms, std::chrono::duration<double, std::milli>(w1 - w0).count(),
                    std::chrono::duration<double, std::milli>(w2 - w0).count());
        }
    } trace_end{g, trace, w0, reqs.size(), super_steps, w0};
    for (int step = 0; step < super_steps && !reqs.empty(); step++) {
        int rc = ensure_staging(ctx, count, 0);
        if (rc) return rc;
        CU(cudaMemsetAsync(ctx->d_fail, 0, count * sizeof(int), ctx->stream));
        ctx->ext_skip = g->S.skip;
        rc = run_batch(ctx, reqs, g->S.theta, 1, g->S.nlml, g->S.grad, g->S.status, nullptr, nullptr, 0);
        if (rc == MEDGP_OK && !ctx->retry_on_device)
            rc = host_retry_rounds(ctx, reqs, count, g->S.theta, 1, g->S.nlml, g->S.grad, g->S.status, nullptr, nullptr);
        ctx->ext_skip = nullptr;
        if (rc) return rc;
        rc = release_desc_slot(ctx);
        if (rc) return rc;
        rc = scg_advance(g);
        if (rc) return rc;
    }
    if (trace) {
        cudaEventRecord(tr_b, ctx->stream);
        trace_end.w1 = std::chrono::steady_clock::now();
    }
    return scg_count_active(g, active_left);
}

MEDGP_API int medgp_cuda_scg_result(medgp_scg *g, double *theta_best, double *loss, int *evals)
{
    if (!g || !g->started) return MEDGP_ERR_ARG;
    medgp_ctx *ctx = g->ctx;
    cudaSetDevice(ctx->device);
    const int count = g->S.count;
    const size_t P = ctx->md.P;
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<ScgScalars> sc(count);
    CU(cudaMemcpy(sc.data(), g->S.sc, (size_t)count * sizeof(ScgScalars), cudaMemcpyDeviceToHost));
    for (int b = 0; b < count; b++) {
        if (loss) loss[b] = sc[b].fX;
        if (evals) evals[b] = sc[b].n_eval;
    }
    if (theta_best)
        CU(cudaMemcpy2D(theta_best, P * 8, g->S.vec + SCG_X * P, SCG_NVEC * P * 8, P * 8, count, cudaMemcpyDeviceToHost));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_scg_points(medgp_scg *g, double *theta, int *wants)
{
    if (!g || !g->started) return MEDGP_ERR_ARG;
    medgp_ctx *ctx = g->ctx;
    cudaSetDevice(ctx->device);
    const int count = g->S.count;
    CU(cudaStreamSynchronize(ctx->stream));
    if (theta) CU(cudaMemcpy(theta, g->S.theta, (size_t)count * ctx->md.P * 8, cudaMemcpyDeviceToHost));
    if (wants) {
        std::vector<int> skip(count);
        CU(cudaMemcpy(skip.data(), g->S.skip, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost));
        for (int b = 0; b < count; b++) wants[b] = skip[b] ? 0 : 1;
    }
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_scg_feed(medgp_scg *g, const double *f, const double *grad, const int *ok)
{
    if (!g || !g->started || !f || !grad || !ok) return MEDGP_ERR_ARG;
    medgp_ctx *ctx = g->ctx;
    cudaSetDevice(ctx->device);
    const int count = g->S.count;
    const size_t P = ctx->md.P;
    std::vector<int> status(count);
    for (int b = 0; b < count; b++) status[b] = ok[b] ? 0 : -1;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(g->S.nlml, f, (size_t)count * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g->S.grad, grad, (size_t)count * P * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g->S.status, status.data(), (size_t)count * sizeof(int), cudaMemcpyHostToDevice));
    int rc = scg_advance(g);
    if (rc) return rc;
    return scg_count_active(g, nullptr);
}

#ifdef MEDGP_X_TRACE
MEDGP_API int medgp_cuda_debug_flow_trace(unsigned long long *out /* 64 x 16 */)
{
    return cudaMemcpyFromSymbol(out, g_flow_trace, sizeof(unsigned long long) * MEDGP_FLOW_TMAX * 16) == cudaSuccess ? 0 : 1;
}
#endif

// ======================================================================= mode-kernel KDE
MEDGP_API int medgp_cuda_kde_mode(medgp_ctx *ctx, int n_sets, const int *offsets, const double *data,
                                  const double *bandwidth, double *mode, double *density)
{
    if (!ctx || n_sets < 0 || !offsets || !data || !bandwidth || !mode) {
        if (ctx) ctx->err = "kde_mode: bad argument";
        return MEDGP_ERR_ARG;
    }
    if (n_sets == 0) return MEDGP_OK;
    int n_max = 0;
    for (int s = 0; s < n_sets; s++) {
        const int n = offsets[s + 1] - offsets[s];
        if (n < 1 || !(bandwidth[s] > 0.0)) {
            ctx->err = "kde_mode: every set needs at least one value and a positive bandwidth";
            return MEDGP_ERR_ARG;
        }
        n_max = std::max(n_max, n);
    }
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const size_t total = (size_t)offsets[n_sets] - (size_t)offsets[0];
    if (offsets[0] != 0) { ctx->err = "kde_mode: offsets must start at 0"; return MEDGP_ERR_ARG; }
    double *d_data = nullptr, *d_bw = nullptr, *d_dens = nullptr, *d_mode = nullptr;
    int *d_off = nullptr;
    auto cleanup = [&]() { cudaFree(d_data); cudaFree(d_bw); cudaFree(d_dens); cudaFree(d_mode); cudaFree(d_off); };
    if (cudaMalloc(&d_data, total * 8) != cudaSuccess || cudaMalloc(&d_dens, total * 8) != cudaSuccess ||
        cudaMalloc(&d_bw, (size_t)n_sets * 8) != cudaSuccess || cudaMalloc(&d_mode, (size_t)n_sets * 8) != cudaSuccess ||
        cudaMalloc(&d_off, (size_t)(n_sets + 1) * sizeof(int)) != cudaSuccess) {
        cleanup();
        ctx->err = "kde_mode: out of device memory";
        return MEDGP_ERR_NOMEM;
    }
    cudaMemcpyAsync(d_data, data, total * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_bw, bandwidth, (size_t)n_sets * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_off, offsets, (size_t)(n_sets + 1) * sizeof(int), cudaMemcpyHostToDevice, st);
    for (int s0 = 0; s0 < n_sets; s0 += 65535) {  // grid.y limit
        const int ns = std::min(65535, n_sets - s0);
        k_kde_density<<<dim3((n_max + MEDGP_KDE_THREADS - 1) / MEDGP_KDE_THREADS, ns), MEDGP_KDE_THREADS, 0, st>>>(
            d_data, d_off + s0, d_bw + s0, d_dens);
    }
    k_kde_mode<<<n_sets, MEDGP_KDE_THREADS, 0, st>>>(d_data, d_off, d_dens, d_mode);
    cudaMemcpyAsync(mode, d_mode, (size_t)n_sets * 8, cudaMemcpyDeviceToHost, st);
    if (density) cudaMemcpyAsync(density, d_dens, total * 8, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    const cudaError_t e2 = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess || e2 != cudaSuccess) {
        ctx->err = std::string("kde_mode: ") + cudaGetErrorString(e != cudaSuccess ? e : e2);
        return MEDGP_ERR_CUDA;
    }
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_debug_force_fail(medgp_ctx *ctx, int attempts)
{
    if (!ctx || attempts < 0) return MEDGP_ERR_ARG;
    ctx->force_fail = attempts;
    return MEDGP_OK;
}

// The reference's out-parameters of c_inference::compute_nlml (chol_alpha, chol_factor_inv:
// inference/c_inference_exact.cpp:124-143), for callers that keep the reference's own
// GP_Regression::predict: alpha = K^-1 y and the row-major lower-triangular L^-1 (strict upper
// part zero), as floats, in the point order of the series -- which must have been uploaded with
// MEDGP_ORDER_GIVEN so that this is the caller's order.
MEDGP_API int medgp_cuda_export_factors(medgp_ctx *ctx, int series_id, const double *theta, float *alpha,
                                        float *Linv, double *nlml, int *status)
{
    if (!ctx || !ctx->model_set || !theta || !alpha || !Linv || !nlml || !status) {
        if (ctx) ctx->err = "export_factors: bad argument";
        return MEDGP_ERR_ARG;
    }
    cudaSetDevice(ctx->device);
    int rc = check_series_ids(ctx, 1, &series_id);
    if (rc) return rc;
    const Series &s = ctx->series[series_id];
    if (!s.given_order) {
        ctx->err = "export_factors: the series must be uploaded with MEDGP_ORDER_GIVEN (the factor refers to the point order)";
        return MEDGP_ERR_ARG;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    rc = ensure_staging(ctx, 1, 0);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const size_t n = s.n, np = s.npad, T = s.T;
    CU(cudaMemcpyAsync(ctx->d_theta, theta, (size_t)ctx->md.P * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(ctx->d_fail, 0, sizeof(int), st));
    std::vector<Request> one = {{series_id, 0, 0, 0, 0}};
    rc = run_batch(ctx, one, ctx->d_theta, 4, ctx->d_out, nullptr, ctx->d_status, nullptr, nullptr, 0);
    if (rc) return rc;
    if (!ctx->retry_on_device) {
        rc = host_retry_rounds(ctx, one, 1, ctx->d_theta, 4, ctx->d_out, nullptr, ctx->d_status, nullptr, nullptr);
        if (rc) return rc;
    }
    rc = release_desc_slot(ctx);
    if (rc) return rc;
    // the single evaluation's buffers are the first arena allocations (see run_batch's carving)
    const EvalDesc &e = ctx->h_descs[0];
    std::vector<double> hT(T * T * kTileElems), hX(T * kTileElems), ha(np);
    CU(cudaMemcpyAsync(hT.data(), e.M, hT.size() * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hX.data(), e.dinv, hX.size() * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ha.data(), e.alpha, np * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_out, ctx->d_out, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    resolve_marks(ctx);
    *status = ctx->h_status[0];
    *nlml = ctx->h_out[0];
    if (*status < 0) return MEDGP_OK;
    for (size_t i = 0; i < n; i++) {
        alpha[i] = (float)ha[i];
        float *row = Linv + i * n;
        for (size_t j = 0; j < n; j++) {
            double v = 0.0;
            if (j <= i) {
                // L^-1(i, j) = U(j, i): off-diagonal tiles of U live strictly above the diagonal
                // of M; the diagonal tiles of L^-1 are X_kk (dinv)
                v = ((i >> 6) == (j >> 6)) ? hX[(i >> 6) * (size_t)kTileElems + (j & 63) * MEDGP_SLD + (i & 63)]
                                           : hT[elem_off((int)T, (int)j, (int)i)];
            }
            row[j] = (float)v;
        }
    }
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_profile(medgp_ctx *ctx, int enable)
{
    if (!ctx) return MEDGP_ERR_ARG;
    ctx->profile = enable != 0;
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_stage_times(medgp_ctx *ctx, medgp_stage_times *out, int reset)
{
    if (!ctx || !out) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    resolve_marks(ctx);
    *out = ctx->times;
    if (reset) ctx->times = medgp_stage_times{};
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_malloc(medgp_ctx *ctx, size_t bytes, void **d_ptr)
{
    if (!ctx || !d_ptr) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaMalloc(d_ptr, bytes));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_free(medgp_ctx *ctx, void *d_ptr)
{
    if (!ctx) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaFree(d_ptr));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_memcpy_h2d(medgp_ctx *ctx, void *d_dst, const void *src, size_t bytes)
{
    if (!ctx) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MEDGP_OK;
}

MEDGP_API int medgp_cuda_memcpy_d2h(medgp_ctx *ctx, void *dst, const void *d_src, size_t bytes)
{
    if (!ctx) return MEDGP_ERR_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MEDGP_OK;
}

MEDGP_API void *medgp_cuda_stream(medgp_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
