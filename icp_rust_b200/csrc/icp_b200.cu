// icp_b200.cu — host side of libicp_b200.so: the C ABI of include/icp_b200.h over the sm_100a
// kernels in this directory.  No CPU compute path exists here: every entry point either launches
// CUDA kernels or returns an error.
#include "../../include/icp_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "icp_instances.h"
#include "icp_stats_kernel.cuh"
#include "icp_target.cuh"

using namespace icp;

namespace {

// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// kernel instances live in their own translation units (icp_fused_inst.cu, list in the
// Makefile) so that they compile in parallel
const Instance* const kInstances[] = {&kFusedInstance0, &kFusedInstance1, &kFusedInstance2, &kFusedInstance3,
                                      &kFusedInstance4, &kFusedInstance5, &kFusedInstance6, &kFusedInstance7,
                                      &kFusedInstance8, &kFusedInstance9, &kFusedInstance10, &kFusedInstance11,
                                      &kFusedInstance12, &kFusedInstance13};
constexpr int kNumInstances = (int)(sizeof(kInstances) / sizeof(kInstances[0]));
constexpr int kMaxChunks = 64;
// device counter words of a context: per-chunk work counters | error word | work counter of the
// ordered re-run | count of flagged pairs
// host-pointer calls: one work counter per chunk, the call's status word, and per chunk the work counter and the
// count of the guard's re-run (a chunk's flagged pairs are listed at redo_list + the chunk's first pair)
constexpr int kCtrWork = 0, kCtrErr = kMaxChunks, kCtrRedoWork = kMaxChunks + 1, kCtrRedoCount = 2 * kMaxChunks + 1;
constexpr int kCtrWords = 3 * kMaxChunks + 1;
// Counters of the device-pointer calls: two alternating sets (call parity) the kernels clean up
// themselves (FusedParams::fin_counter), so back-to-back calls need no memset between them and may
// overlap at their boundary (ICP_B200_FLAG_INDEPENDENT).  Words of set s at s * kDevSetWords.
constexpr int kDevWork = 0, kDevFin = 1, kDevRedoWork = 2, kDevRedoFin = 3, kDevRedoCount = 4, kDevRedoSeen = 5;
constexpr int kDevSetWords = 8, kDevErr = 2 * kDevSetWords, kDevWords = kDevErr + 1;
constexpr int kOverlapRedoGrid = 8;   // CTAs of the re-run launch of an overlapping call (they wait on an SM each)

int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

}  // namespace

namespace { struct HostPending; }

struct icp_b200_context {
    HostPending* pending = nullptr;   // a submitted host-pointer batch that has not been waited for
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaStream_t redo_stream = nullptr;   // host-pointer path: the guard's re-runs, chunk by chunk, beside the later chunks
    cudaEvent_t ev_cnt[kMaxChunks];       // chunk c's count of flagged pairs has landed in h_redo[c]
    uint32_t* h_redo = nullptr;           // pinned, kMaxChunks words
    cudaEvent_t ev[kMaxChunks][2];
    cudaEvent_t ev_ready = nullptr;
    cudaEvent_t ev_redo[2] = {nullptr, nullptr};
    bool events_ok = false;
    DevBuf src, dst, src_off, dst_off, init_T, out_T, out_inner, out_corr, out_hash, out_status, out_cost;
    DevBuf counters;       // kCtrWords words (layout above)
    DevBuf workspace[2];
    DevBuf redo_list, redo_workspace;   // flagged pairs of the call in flight / slabs of their ordered re-run
    // device-pointer calls: per-parity state, so that a call may start while the previous one drains
    DevBuf dev_counters, dev_workspace[2], dev_redo_list[2], dev_redo_workspace[2];
    uint64_t dev_calls = 0;
    bool dev_dirty = true;  // the counters need a memset (first call, or a call that failed midway)
    int blocking_sync = 0;  // ICP_B200_BLOCKING_SYNC=1: the host-pointer path sleeps on a blocking event instead of spinning in cudaStreamSynchronize
    cudaEvent_t ev_block = nullptr;
    int pdl = 1;            // ICP_B200_PDL=0: never overlap calls (ICP_B200_FLAG_INDEPENDENT is ignored)
    int fixed_point = 1;    // ICP_B200_FIXED_POINT=0: compute every outer iteration, also the ones that provably repeat
    int guard = 1;          // ICP_B200_GUARD=0: no knife-edge guard (as ICP_B200_FLAG_NO_GUARD on every call)
    double guard_scale = 1.0;   // ICP_B200_GUARD_SCALE: multiplies the guard bands (tests force re-runs with it)
    DevBuf stats_in, stats_off, stats_out;
    icp_b200_call_info info{};
    std::string last_error;
    int cell_budget = 4096;
    int force_instance = -1;
    int use_lean = 1;       // ICP_B200_LEAN=0: warp-tree batches on the icp_fused.cuh schedule instead of icp_lean.cuh
    int stream_cells_per_point = 0;   // ICP_B200_STREAM_CELLS_PER_POINT: cell budget of the streaming kernel's per-pair grids (0: 8 in 2-D, 4 in 3-D)
    int single_ordered = 0;     // ICP_B200_SINGLE_ORDERED=1: the single-pair entries use ordered sums
    int stream_nn_f32 = 1;      // ICP_B200_STREAM_NN_F32=0: the streaming kernel's 3-D searches rank every candidate in binary64
    int stream_nn_group = 0;    // ICP_B200_STREAM_NN_GROUP=1: a group of lanes per query in the streaming kernel's search
    int force_stream = 0;   // ICP_B200_FORCE_STREAM=1: route every warp-tree batch to the streaming kernel (tests)
};

struct icp_b200_target {
    icp_b200_context* ctx = nullptr;
    uint32_t dim = 2;
    uint32_t m = 0;
    GridGeom g{};
    DevBuf pts, x, y, z, idx, cstart, cnt, slot, keys;
    DevBuf pf, seed;          // 2-D fast search (icp_target_nn.cuh); seed_ptr null when it is off
    const uint32_t* seed_ptr = nullptr;
    std::vector<double> host_pts;             // the point set (AoS), kept for icp_b200_target_update
    double bbox[6] = {0, 0, 0, 0, 0, 0};      // bounding box the geometry was built for (min xyz, max xyz)
    uint32_t m_built = 0;                     // point count the geometry was built for
    DevBuf idx_prev, remap, where, seed2;     // scratch of the incremental update
    DevBuf queries, out_idx;
};

namespace {

int fail(icp_b200_context* ctx, int status, const char* what, cudaError_t e = cudaSuccess)
{
    if (ctx) {
        ctx->last_error = what ? what : "";
        if (e != cudaSuccess) {
            ctx->last_error += ": ";
            ctx->last_error += cudaGetErrorString(e);
        }
    }
    return status;
}

#define ICP_CUDA(ctx, call)                                                        \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            cudaGetLastError();                                                    \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? ICP_B200_ERR_OUT_OF_MEMORY \
                                                              : ICP_B200_ERR_CUDA, \
                        #call, e__);                                               \
        }                                                                          \
    } while (0)

int pair_status_to_error(int st)
{
    switch (st) {
        case kPairOk: return ICP_B200_OK;
        case kPairEmptyTarget: return ICP_B200_ERR_EMPTY_TARGET;
        case kPairTooLarge: return ICP_B200_ERR_TOO_LARGE;
        case kPairNonFinite: return ICP_B200_ERR_NON_FINITE;
        default: return ICP_B200_ERR_CUDA;
    }
}

struct LaunchPlan {
    const Instance* inst = nullptr;   // fused per-pair kernel ...
    FusedLayout layout{};
    int mcap = 0;
    int grid = 0;
    FusedKernel kernel = nullptr;
    bool lean = false;                // icp_lean.cuh schedule of the warp-tree mode
    bool stream = false;              // ... or the streaming kernel for large scans
    StreamKernel skernel = nullptr;
    StreamLayout slayout{};
    bool stream_ordered = false;
    uint32_t n_cap = 0, m_cap = 0, c_cap = 0, cells_per_point = 8;
    int threads = 0;
    uint32_t smem = 0;
    uint64_t ws_stride = 0;
};

// Chooses the kernel instance for the largest scan of the batch and sizes its launch.
int plan_launch(icp_b200_context* ctx, uint32_t dim, bool shared_target, bool ordered, uint32_t max_n, uint32_t max_m,
                uint32_t pairs, LaunchPlan* plan)
{
    const Instance* inst = nullptr;
    auto lean_of = [&](const Instance* c) { return !ordered && ctx->use_lean && c->lean[dim - 2][shared_target ? 1 : 0] != nullptr; };
    auto fits = [&](const Instance* c) { return c->ordered == ordered && (uint32_t)c->capacity(lean_of(c)) >= max_n; };
    if (ctx->force_instance >= 0 && ctx->force_instance < kNumInstances && fits(kInstances[ctx->force_instance]))
        inst = kInstances[ctx->force_instance];
    auto shape = [&](const Instance* c) {      // fills the plan for instance c; false when it does not fit one SM
        plan->inst = c;
        plan->mcap = (int)((max_m + 7u) & ~7u);
        plan->lean = lean_of(c);
        plan->layout = fused_layout((int)dim, c->capacity(plan->lean), plan->mcap, ctx->cell_budget, shared_target, ordered);
        if (plan->lean)
            plan->layout.total = lean_layout((int)dim, c->capacity(true), plan->mcap, ctx->cell_budget, shared_target).total;
        if (!shared_target && max_m > 65528u) return false;                      // u16 indices of the shared-memory grid
        return (int)plan->layout.total <= ctx->max_smem_optin;
    };
    if (inst && !shape(inst)) inst = nullptr;
    // Fewer pairs than SMs (the reference's own call pattern is ONE pair per frame, examples/scan2d.rs:87-88): every
    // pair has an SM to itself and latency is what counts — the widest instance with dedicated service warps first
    // (384 threads, 0.59 ms against 0.69 ms for the bundled 666-point pair), then the wide ones for larger scans.
    if (!inst && !ordered && pairs < (uint32_t)ctx->sm_count) {
        static const int kLatencyOrder[] = {3, 5, 7, 6};
        for (int i : kLatencyOrder)
            if (!inst && i < kNumInstances && fits(kInstances[i]) && shape(kInstances[i])) inst = kInstances[i];
    }
    // Ordered sums, fewer pairs than SMs: the 256-thread ordered instance takes 3.8 ms for a lone 1080-point pair
    // (1.11 ms for the bundled 666-point pair), the 128-thread one 4.9 ms (1.40 ms).  Not for the guard's re-run of a
    // large batch (plan_redo plans for the batch size): a 256-thread ordered CTA needs a whole SM's registers and
    // would wait for one to drain, while the 128-thread one fits the hole a single retiring tree-mode CTA leaves.
    if (!inst && ordered && pairs < (uint32_t)ctx->sm_count) {
        static const int kOrderedLatency[] = {12, 13};
        for (int i : kOrderedLatency)
            if (!inst && i < kNumInstances && fits(kInstances[i]) && shape(kInstances[i])) inst = kInstances[i];
    }
    for (int i = 0; !inst && i < kNumInstances; ++i)     // the first instance that holds the scan AND fits an SM
        if (fits(kInstances[i]) && shape(kInstances[i])) inst = kInstances[i];
    if (ctx->force_stream && !ordered) inst = nullptr;
    if (!inst) {
        // large scans: streaming kernel (warp-tree sums, or the tiled reference-order chain)
        if (max_n > 65535u) return fail(ctx, ICP_B200_ERR_TOO_LARGE, "source scan larger than 65535 points");
        plan->inst = nullptr;
        plan->stream = true;
        plan->stream_ordered = ordered;
        plan->n_cap = std::max<uint32_t>(8u, (max_n + 7u) & ~7u);
        plan->m_cap = shared_target ? 0u : std::max<uint32_t>(8u, (max_m + 7u) & ~7u);
        // cells per target point: measured on config 4 after the fixed-point exit (1024 pairs): 2 / 4 / 8 / 16
        // cells -> 6.25k / 6.69k / 6.05k / 5.41k pairs/s
        plan->cells_per_point = ctx->stream_cells_per_point > 0 ? (uint32_t)ctx->stream_cells_per_point : (dim == 3 ? 4u : 8u);
        plan->c_cap = shared_target ? 0u : std::max<uint32_t>(64u, plan->cells_per_point * plan->m_cap);
        plan->slayout = stream_layout((int)dim, plan->n_cap, plan->m_cap, plan->c_cap, shared_target, ordered);
        plan->ws_stride = plan->slayout.total;
        plan->skernel = kStreamInstance.k[dim - 2][shared_target ? 1 : 0];
        plan->threads = kStreamInstance.threads;
        plan->smem = stream_smem_bytes(ordered);
        int occ = 0;
        ICP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, plan->skernel, plan->threads, plan->smem));
        if (occ < 1) return fail(ctx, ICP_B200_ERR_CUDA, "streaming kernel cannot be resident (occupancy 0)");
        plan->grid = (int)std::max<long long>(1, std::min<long long>((long long)occ * ctx->sm_count, pairs));
        return ICP_B200_OK;
    }
    plan->kernel = plan->lean ? inst->lean[dim - 2][shared_target ? 1 : 0] : inst->k[dim - 2][shared_target ? 1 : 0];
    plan->threads = inst->threads;
    plan->smem = plan->layout.total;
    plan->ws_stride = ((uint64_t)plan->layout.nn_bytes + 255u) & ~255ull;
    ICP_CUDA(ctx, cudaFuncSetAttribute(plan->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->layout.total));
    int occ = 0;
    ICP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, plan->kernel, inst->threads, plan->layout.total));
    if (occ < 1) return fail(ctx, ICP_B200_ERR_CUDA, "fused kernel cannot be resident (occupancy 0)");
    long long g = (long long)occ * ctx->sm_count;
    plan->grid = (int)std::max<long long>(1, std::min<long long>(g, pairs));
    return ICP_B200_OK;
}


int validate_batch(icp_b200_context* ctx, const icp_b200_batch* b, bool shared_target)
{
    if (!ctx || !b) return ICP_B200_ERR_INVALID_ARGUMENT;
    if (b->dim != 2 && b->dim != 3) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "dim must be 2 or 3");
    if (b->batch == 0) return ICP_B200_OK;
    if (!b->src_off || !b->init_T || !b->out_T) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null src_off / init_T / out_T");
    if (!shared_target && !b->dst_off) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null dst_off");
    if (b->corr_mode < 0 || b->corr_mode > 2) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "bad corr_mode");
    if (b->flags & ~(uint32_t)(ICP_B200_FLAG_ORDERED_SUMS | ICP_B200_FLAG_SHARED_SOURCE | ICP_B200_FLAG_NO_GUARD |
                               ICP_B200_FLAG_INDEPENDENT))
        return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "unknown flag bits");
    if (b->corr_mode != ICP_B200_CORR_NONE && !b->out_corr) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "corr_mode set but out_corr is null");
    return ICP_B200_OK;
}

void fill_params(icp_b200_context* ctx, const icp_b200_batch& b, const icp_b200_target* tgt, FusedParams* out)
{
    FusedParams& p = *out;
    p.src = b.src; p.src_off = b.src_off; p.dst = b.dst; p.dst_off = b.dst_off; p.init_T = b.init_T;
    p.out_T = b.out_T; p.out_inner = b.out_inner; p.out_corr = b.out_corr; p.out_hash = b.out_corr_hash;
    p.out_status = b.out_status; p.out_cost = b.out_cost;
    p.shared_src = (b.flags & ICP_B200_FLAG_SHARED_SOURCE) ? 1u : 0u;
    p.fixed_point = ctx->fixed_point ? 1u : 0u;
    p.out_error = ctx->counters.as<int32_t>() + kCtrErr;
    p.max_iter = b.max_iter; p.corr_mode = b.corr_mode;
    p.cell_budget = ctx->cell_budget;
    p.sm_count = (uint32_t)ctx->sm_count;
    if (tgt) {
        p.target.x = tgt->x.as<double>(); p.target.y = tgt->y.as<double>(); p.target.z = tgt->z.as<double>();
        p.target.idx = tgt->idx.as<uint32_t>(); p.target.cstart = tgt->cstart.as<uint32_t>();
        p.target.g = tgt->g; p.target.m = tgt->m;
        p.target.seed = tgt->seed_ptr; p.target.pf = tgt->seed_ptr ? tgt->pf.as<float2>() : nullptr;
    }
}

// `overlap`: launch with the programmatic-serialization attribute — the kernel may become resident
// while the kernel before it on `stream` drains (icp_fused.cuh fused_launch_entry has the rules).
int launch_plan(icp_b200_context* ctx, const LaunchPlan& pl, const FusedParams& fp, int grid, cudaStream_t stream,
                bool overlap = false)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)pl.threads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1u : 0u;
    if (pl.stream) {
        StreamParams sp{};
        sp.f = fp; sp.n_cap = pl.n_cap; sp.m_cap = pl.m_cap; sp.c_cap = pl.c_cap;
        sp.cells_per_point = pl.cells_per_point;
        sp.ordered = pl.stream_ordered ? 1u : 0u;
        sp.nn_group = (uint32_t)ctx->stream_nn_group;
        sp.nn_f32 = (uint32_t)ctx->stream_nn_f32;
        ICP_CUDA(ctx, cudaLaunchKernelEx(&cfg, pl.skernel, sp));
    } else {
        ICP_CUDA(ctx, cudaLaunchKernelEx(&cfg, pl.kernel, fp));
    }
    ICP_CUDA(ctx, cudaGetLastError());
    ctx->info.kernel_launches += 1;
    return ICP_B200_OK;
}

// Device-pointer launch of pairs [begin, end) on `stream`.  `guarded` (tree mode with the knife-edge
// guard): pairs with a decision inside a guard band are appended to the call's redo list on the
// device (one list and one count per call, shared by its chunks).
// `dev_set` >= 0: a device-pointer call — the counters, workspace and flagged-pair list of that
// parity, cleaned up by the kernels; `overlap`: the launch may start while its predecessor drains.
int launch_range(icp_b200_context* ctx, const icp_b200_batch& b, const icp_b200_target* tgt, const LaunchPlan& plan,
                 bool guarded, uint32_t begin, uint32_t end, int chunk_slot, int ws_slot, cudaStream_t stream,
                 int dev_set = -1, bool overlap = false)
{
    FusedParams p{};
    fill_params(ctx, b, tgt, &p);
    uint32_t* dev = dev_set >= 0 ? ctx->dev_counters.as<uint32_t>() + dev_set * kDevSetWords : nullptr;
    p.work_counter = dev ? dev + kDevWork : ctx->counters.as<uint32_t>() + kCtrWork + chunk_slot;
    p.workspace = dev ? ctx->dev_workspace[dev_set].as<unsigned char>() : ctx->workspace[ws_slot].as<unsigned char>();
    if (dev) {
        p.fin_counter = dev + kDevFin;
        p.out_error = ctx->dev_counters.as<int32_t>() + kDevErr;
    }
    p.ws_stride = plan.ws_stride;
    p.pair_begin = begin; p.pair_end = end;
    p.mcap = plan.mcap;
    p.ncap = plan.stream ? (int)plan.n_cap : plan.inst->capacity(plan.lean);
    if (guarded) {
        p.redo_list = dev ? ctx->dev_redo_list[dev_set].as<uint32_t>() : ctx->redo_list.as<uint32_t>() + begin;
        p.redo_count = dev ? dev + kDevRedoCount : ctx->counters.as<uint32_t>() + kCtrRedoCount + chunk_slot;
        p.guard.det = kGuardDet * ctx->guard_scale;
        p.guard.delta = kGuardDelta * ctx->guard_scale;
        p.guard.err = kGuardErr * ctx->guard_scale;
    }
    const int grid = (int)std::min<uint32_t>((uint32_t)plan.grid, end - begin);
    int rc = launch_plan(ctx, plan, p, grid, stream, overlap);
    if (rc != ICP_B200_OK) return rc;
    ctx->info.grid_ctas = (uint32_t)grid;
    ctx->info.block_threads = (uint32_t)plan.threads;
    ctx->info.smem_bytes = plan.smem;
    return ICP_B200_OK;
}

// The ordered re-run of the pairs the call's tree-mode launches flagged: one launch behind them on
// `stream`, the flagged pairs spread over its CTAs (their indices and count never leave the device).
// A lone 1080-point pair takes 4-5 ms in reference order (2n dependent additions per evaluation).  Device-pointer
// calls issue the re-run once per call (it overlaps the next call).  Host-pointer calls learn each chunk's count of
// flagged pairs as the chunk ends and re-run exactly those chunks on a stream of their own, beside the chunks
// still to come (host_wait) — a flagged pair costs the call nothing unless it sits in the last chunk.
int launch_redo(icp_b200_context* ctx, const icp_b200_batch& b, const icp_b200_target* tgt, const LaunchPlan& redo,
                cudaStream_t stream, int dev_set = -1, bool overlap = false, int chunk = 0, uint32_t chunk_begin = 0,
                uint32_t flagged = 0)
{
    FusedParams r{};
    fill_params(ctx, b, tgt, &r);
    if (dev_set >= 0) {
        uint32_t* dev = ctx->dev_counters.as<uint32_t>() + dev_set * kDevSetWords;
        r.pair_list = ctx->dev_redo_list[dev_set].as<uint32_t>();
        r.pair_count = dev + kDevRedoCount;
        r.work_counter = dev + kDevRedoWork;
        r.fin_counter = dev + kDevRedoFin;
        r.redo_seen = dev + kDevRedoSeen;
        r.out_error = ctx->dev_counters.as<int32_t>() + kDevErr;
        r.workspace = ctx->dev_redo_workspace[dev_set].as<unsigned char>();
    } else {
        // host-pointer call: the re-run of ONE chunk's flagged pairs (their count is known to the host by now)
        r.pair_list = ctx->redo_list.as<uint32_t>() + chunk_begin;
        r.pair_count = ctx->counters.as<uint32_t>() + kCtrRedoCount + chunk;
        r.work_counter = ctx->counters.as<uint32_t>() + kCtrRedoWork + chunk;
        r.workspace = ctx->redo_workspace.as<unsigned char>();
    }
    r.ws_stride = redo.ws_stride;
    r.pair_begin = 0; r.pair_end = b.batch;
    r.mcap = redo.mcap;
    r.ncap = redo.stream ? (int)redo.n_cap : redo.inst->capacity(false);
    uint32_t grid = std::min<uint32_t>((uint32_t)redo.grid, b.batch);
    if (dev_set < 0) grid = std::max<uint32_t>(1u, std::min<uint32_t>(grid, flagged));
    return launch_plan(ctx, redo, r, (int)grid, stream, overlap);
}

// The ordered plan that re-runs flagged pairs, when this call wants the guard.
int plan_redo(icp_b200_context* ctx, const icp_b200_batch* b, bool shared, uint32_t max_n, uint32_t max_m, LaunchPlan* redo,
              bool* use, int dev_set = -1, bool overlap = false)
{
    *use = false;
    if ((b->flags & (ICP_B200_FLAG_ORDERED_SUMS | ICP_B200_FLAG_NO_GUARD)) || !ctx->guard) return ICP_B200_OK;
    int rc = plan_launch(ctx, b->dim, shared, true, max_n, max_m, b->batch, redo);
    if (rc != ICP_B200_OK) return rc;
    // a handful of pairs: one CTA per SM is plenty; an overlapping call keeps it to a few CTAs, which
    // sit on their SMs until the main launch has finished
    redo->grid = std::min(redo->grid, overlap ? kOverlapRedoGrid : ctx->sm_count);
    DevBuf& list = dev_set >= 0 ? ctx->dev_redo_list[dev_set] : ctx->redo_list;
    DevBuf& slabs = dev_set >= 0 ? ctx->dev_redo_workspace[dev_set] : ctx->redo_workspace;
    ICP_CUDA(ctx, list.ensure((size_t)b->batch * 4));
    ICP_CUDA(ctx, slabs.ensure((size_t)redo->grid * redo->ws_stride + 256));
    *use = true;
    return ICP_B200_OK;
}

int ensure_events(icp_b200_context* ctx)
{
    if (ctx->events_ok) return ICP_B200_OK;
    for (int i = 0; i < kMaxChunks; ++i)
        for (int j = 0; j < 2; ++j) ICP_CUDA(ctx, cudaEventCreate(&ctx->ev[i][j]));
    ICP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    for (int j = 0; j < 2; ++j) ICP_CUDA(ctx, cudaEventCreate(&ctx->ev_redo[j]));
    ICP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
    for (int i = 0; i < kMaxChunks; ++i)
        ICP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_cnt[i], cudaEventDisableTiming | (ctx->blocking_sync ? cudaEventBlockingSync : 0)));
    ICP_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_redo), kMaxChunks * sizeof(uint32_t), cudaHostAllocDefault));
    ctx->events_ok = true;
    return ICP_B200_OK;
}

// Wait of the host-pointer path for one of its streams.  Many ranks on few host cores: a thread that spins in
// cudaStreamSynchronize for the 20+ ms of a batch takes a core from the ranks beside it; the blocking form sleeps.
int host_sync(icp_b200_context* ctx, cudaStream_t st)
{
    if (ctx->blocking_sync) {
        ICP_CUDA(ctx, cudaEventRecord(ctx->ev_block, st));
        ICP_CUDA(ctx, cudaEventSynchronize(ctx->ev_block));
        return ICP_B200_OK;
    }
    ICP_CUDA(ctx, cudaStreamSynchronize(st));
    return ICP_B200_OK;
}

int estimate_device_impl(icp_b200_context* ctx, const icp_b200_target* tgt, const icp_b200_batch* b, void* stream_v,
                         int synchronize)
{
    const bool shared = tgt != nullptr;
    int rc = validate_batch(ctx, b, shared);
    if (rc != ICP_B200_OK || b->batch == 0) return rc;
    if (ctx->pending) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "a submitted batch is waiting for icp_b200_estimate_batch_wait");
    ICP_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream[0];
    ctx->info = icp_b200_call_info{};
    rc = ensure_events(ctx);
    if (rc != ICP_B200_OK) return rc;

    uint32_t max_n = b->max_src_points, max_m = shared ? tgt->m : b->max_dst_points;
    if (max_n == 0 || (!shared && max_m == 0)) {
        // sizes unknown: read the offset tables back once (documented in the header)
        std::vector<uint32_t> off(b->batch + 1);
        ICP_CUDA(ctx, cudaMemcpyAsync(off.data(), b->src_off, off.size() * 4, cudaMemcpyDeviceToHost, stream));
        ICP_CUDA(ctx, cudaStreamSynchronize(stream));
        max_n = 0;
        for (uint32_t i = 0; i < b->batch; ++i) max_n = std::max(max_n, off[i + 1] - off[i]);
        if (!shared) {
            ICP_CUDA(ctx, cudaMemcpyAsync(off.data(), b->dst_off, off.size() * 4, cudaMemcpyDeviceToHost, stream));
            ICP_CUDA(ctx, cudaStreamSynchronize(stream));
            max_m = 0;
            for (uint32_t i = 0; i < b->batch; ++i) max_m = std::max(max_m, off[i + 1] - off[i]);
        }
        ctx->info.d2h_bytes += (uint64_t)off.size() * 4 * (shared ? 1 : 2);
    }
    // State of this call's parity; the call two back used it and has completed by the time this
    // call's kernels run (fused_launch_entry).  ICP_B200_FLAG_INDEPENDENT lets the kernels become
    // resident while the previous call's last pairs finish.
    const int set = (int)(ctx->dev_calls & 1u);
    const bool overlap = ctx->pdl && (b->flags & ICP_B200_FLAG_INDEPENDENT) != 0;
    LaunchPlan plan, redo;
    bool use_redo = false;
    rc = plan_launch(ctx, b->dim, shared, (b->flags & ICP_B200_FLAG_ORDERED_SUMS) != 0, max_n, max_m, b->batch, &plan);
    if (rc != ICP_B200_OK) return rc;
    rc = plan_redo(ctx, b, shared, max_n, max_m, &redo, &use_redo, set, overlap);
    if (rc != ICP_B200_OK) return rc;
    ICP_CUDA(ctx, ctx->dev_counters.ensure(kDevWords * 4));
    ICP_CUDA(ctx, ctx->dev_workspace[set].ensure((size_t)plan.grid * plan.ws_stride + 256));
    if (ctx->dev_dirty) ICP_CUDA(ctx, cudaMemsetAsync(ctx->dev_counters.p, 0, kDevWords * 4, stream));
    ctx->dev_dirty = true;   // until every launch of this call is in the stream
    ctx->dev_calls += 1;
    // events only around a synchronous call: a record between two calls would serialise them
    if (synchronize) ICP_CUDA(ctx, cudaEventRecord(ctx->ev[0][0], stream));
    rc = launch_range(ctx, *b, tgt, plan, use_redo, 0, b->batch, 0, 0, stream, set, overlap);
    if (rc != ICP_B200_OK) return rc;
    if (use_redo) {
        rc = launch_redo(ctx, *b, tgt, redo, stream, set, overlap);
        if (rc != ICP_B200_OK) return rc;
    }
    ctx->dev_dirty = false;
    if (synchronize) {
        ICP_CUDA(ctx, cudaEventRecord(ctx->ev[0][1], stream));
        int32_t err = 0;
        uint32_t nredo = 0;
        const uint32_t* dev = ctx->dev_counters.as<uint32_t>() + set * kDevSetWords;
        ICP_CUDA(ctx, cudaMemcpyAsync(&err, ctx->dev_counters.as<int32_t>() + kDevErr, 4, cudaMemcpyDeviceToHost, stream));
        if (use_redo) ICP_CUDA(ctx, cudaMemcpyAsync(&nredo, dev + kDevRedoSeen, 4, cudaMemcpyDeviceToHost, stream));
        ICP_CUDA(ctx, cudaStreamSynchronize(stream));
        ICP_CUDA(ctx, cudaEventElapsedTime(&ctx->info.kernel_ms, ctx->ev[0][0], ctx->ev[0][1]));
        ctx->info.redo_pairs = nredo;
        if (err != 0) {
            // the error word is sticky across asynchronous calls; reading it resets it
            ICP_CUDA(ctx, cudaMemsetAsync(ctx->dev_counters.as<int32_t>() + kDevErr, 0, 4, stream));
            return fail(ctx, pair_status_to_error(err), "at least one scan pair was rejected (see out_status)");
        }
    }
    return ICP_B200_OK;
}

// What a submitted host-pointer batch leaves for its wait: everything the second half of the call needs.
struct HostPending {
    icp_b200_batch hb, db;
    icp_b200_target* tgt = nullptr;
    LaunchPlan redo;
    bool use_redo = false, shared_src = false;
    int used = 0;
    uint64_t ncorr = 0, nsrc = 0;
    uint32_t per = 0;    // pairs per chunk
};

int host_wait(icp_b200_context* ctx);

// Host-pointer path: chunked H2D -> kernel -> D2H pipeline over two streams.  First half: everything is
// ENQUEUED (copies in, kernels, copies out) and nothing waited for; host_wait is the second half.
int host_submit(icp_b200_context* ctx, icp_b200_target* tgt, const icp_b200_batch* hb)
{
    const bool shared = tgt != nullptr;
    int rc = validate_batch(ctx, hb, shared);
    if (rc != ICP_B200_OK) return rc;
    if (ctx->pending) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "a submitted batch is waiting for icp_b200_estimate_batch_wait");
    if (hb->batch == 0) return rc;
    ICP_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->info = icp_b200_call_info{};
    rc = ensure_events(ctx);
    if (rc != ICP_B200_OK) return rc;
    const uint32_t B = hb->batch, dim = hb->dim;
    const bool shared_src = (hb->flags & ICP_B200_FLAG_SHARED_SOURCE) != 0;
    const uint32_t SB = shared_src ? 1u : B;      // entries of the source offset table - 1
    const uint32_t* so = hb->src_off;
    const uint32_t* doff = hb->dst_off;
    uint32_t max_n = 0, max_m = shared ? tgt->m : 0;
    for (uint32_t i = 0; i < B; ++i) {
        if (i < SB) {
            if (so[i + 1] < so[i]) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "src_off not monotone");
            max_n = std::max(max_n, so[i + 1] - so[i]);
        }
        if (!shared) {
            if (doff[i + 1] < doff[i]) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "dst_off not monotone");
            max_m = std::max(max_m, doff[i + 1] - doff[i]);
        }
    }
    const uint64_t nsrc = so[SB], ndst = shared ? 0 : doff[B];
    if ((nsrc && !hb->src) || (ndst && !hb->dst)) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null point array");
    LaunchPlan plan, redo;
    bool use_redo = false;
    rc = plan_launch(ctx, dim, shared, (hb->flags & ICP_B200_FLAG_ORDERED_SUMS) != 0, max_n, max_m, B, &plan);
    if (rc != ICP_B200_OK) return rc;
    rc = plan_redo(ctx, hb, shared, max_n, max_m, &redo, &use_redo);
    if (rc != ICP_B200_OK) return rc;

    const uint64_t ncorr_pts = shared_src ? nsrc * B : nsrc;   // SHARED_SOURCE: laid out as if the scan were repeated
    const uint64_t ncorr = hb->max_iter == 0 ? 0
                         : hb->corr_mode == ICP_B200_CORR_ALL ? ncorr_pts * hb->max_iter
                         : hb->corr_mode == ICP_B200_CORR_LAST ? ncorr_pts : 0;
    ICP_CUDA(ctx, ctx->src.ensure(std::max<uint64_t>(nsrc, 1) * dim * 8));
    ICP_CUDA(ctx, ctx->dst.ensure(std::max<uint64_t>(ndst, 1) * dim * 8));
    ICP_CUDA(ctx, ctx->src_off.ensure((SB + 1) * 4));
    ICP_CUDA(ctx, ctx->dst_off.ensure((B + 1) * 4));
    ICP_CUDA(ctx, ctx->init_T.ensure((size_t)B * 48));
    ICP_CUDA(ctx, ctx->out_T.ensure((size_t)B * 48));
    if (hb->out_inner) ICP_CUDA(ctx, ctx->out_inner.ensure(std::max<size_t>((size_t)B * hb->max_iter, 1) * 4));
    if (ncorr) ICP_CUDA(ctx, ctx->out_corr.ensure(ncorr * 4));
    if (hb->out_corr_hash) ICP_CUDA(ctx, ctx->out_hash.ensure((size_t)B * 8));
    if (hb->out_status) ICP_CUDA(ctx, ctx->out_status.ensure((size_t)B * 4));
    if (hb->out_cost) ICP_CUDA(ctx, ctx->out_cost.ensure((size_t)B * 8));
    ICP_CUDA(ctx, ctx->counters.ensure(kCtrWords * 4));
    for (int s = 0; s < 2; ++s) ICP_CUDA(ctx, ctx->workspace[s].ensure((size_t)plan.grid * plan.ws_stride + 256));

    icp_b200_batch db = *hb;
    db.src = ctx->src.as<double>(); db.dst = ctx->dst.as<double>();
    db.src_off = ctx->src_off.as<uint32_t>(); db.dst_off = ctx->dst_off.as<uint32_t>();
    db.init_T = ctx->init_T.as<double>(); db.out_T = ctx->out_T.as<double>();
    db.out_inner = hb->out_inner ? ctx->out_inner.as<uint32_t>() : nullptr;
    db.out_corr = ncorr ? ctx->out_corr.as<int32_t>() : nullptr;
    db.out_corr_hash = hb->out_corr_hash ? ctx->out_hash.as<uint64_t>() : nullptr;
    db.out_status = hb->out_status ? ctx->out_status.as<int32_t>() : nullptr;
    db.out_cost = hb->out_cost ? ctx->out_cost.as<double>() : nullptr;

    // chunking: enough chunks to overlap copies with compute, each at least one full wave
    int nchunks = (int)std::min<uint32_t>(kMaxChunks, std::max<uint32_t>(1, B / std::max<uint32_t>(1, (uint32_t)plan.grid * 2)));
    nchunks = std::min(nchunks, env_int("ICP_B200_CHUNKS", 8));
    nchunks = std::max(nchunks, 1);
    const uint32_t per = (B + (uint32_t)nchunks - 1) / (uint32_t)nchunks;

    cudaStream_t s0 = ctx->stream[0], s1 = ctx->stream[1];
    ICP_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, kCtrWords * 4, s0));
    ICP_CUDA(ctx, cudaMemcpyAsync(ctx->src_off.p, so, (SB + 1) * 4, cudaMemcpyHostToDevice, s0));
    if (shared_src && nsrc) {
        ICP_CUDA(ctx, cudaMemcpyAsync(ctx->src.p, hb->src, nsrc * dim * 8, cudaMemcpyHostToDevice, s0));
        ctx->info.h2d_bytes += nsrc * dim * 8;
    }
    if (!shared) ICP_CUDA(ctx, cudaMemcpyAsync(ctx->dst_off.p, doff, (B + 1) * 4, cudaMemcpyHostToDevice, s0));
    ICP_CUDA(ctx, cudaMemcpyAsync(ctx->init_T.p, hb->init_T, (size_t)B * 48, cudaMemcpyHostToDevice, s0));
    ICP_CUDA(ctx, cudaEventRecord(ctx->ev_ready, s0));
    ICP_CUDA(ctx, cudaStreamWaitEvent(s1, ctx->ev_ready, 0));
    ctx->info.h2d_bytes += (uint64_t)(SB + 1) * 4 + (shared ? 0 : (uint64_t)(B + 1) * 4) + (uint64_t)B * 48;

    int used = 0;
    for (int c = 0; c < nchunks; ++c) {
        const uint32_t b0 = (uint32_t)c * per, b1 = std::min<uint32_t>(B, b0 + per);
        if (b0 >= b1) break;
        cudaStream_t st = (c & 1) ? s1 : s0;
        // SHARED_SOURCE: the one scan went up with the tables; sp0/sp1 then only place the correspondences
        const uint64_t sp0 = shared_src ? nsrc * b0 : so[b0], sp1 = shared_src ? nsrc * b1 : so[b1];
        // (Measured and dropped: every chunk's inputs issued up front on a copy stream of their own, the kernels
        // gated by events.  All kernels of the next batch are then runnable at once, two callers fall into lockstep
        // and drain together: 169k against 179k pairs/s end to end with two callers, 178k against 179k pipelined.)
        cudaStream_t cs = st;
        if (!shared_src) {
            if (sp1 > sp0)
                ICP_CUDA(ctx, cudaMemcpyAsync(ctx->src.as<double>() + sp0 * dim, hb->src + sp0 * dim, (sp1 - sp0) * dim * 8, cudaMemcpyHostToDevice, cs));
            ctx->info.h2d_bytes += (sp1 - sp0) * dim * 8;
        }
        if (!shared) {
            const uint64_t dp0 = doff[b0], dp1 = doff[b1];
            if (dp1 > dp0)
                ICP_CUDA(ctx, cudaMemcpyAsync(ctx->dst.as<double>() + dp0 * dim, hb->dst + dp0 * dim, (dp1 - dp0) * dim * 8, cudaMemcpyHostToDevice, cs));
            ctx->info.h2d_bytes += (dp1 - dp0) * dim * 8;
        }
        ICP_CUDA(ctx, cudaEventRecord(ctx->ev[c][0], st));
        rc = launch_range(ctx, db, tgt, plan, use_redo, b0, b1, c, c & 1, st);
        if (rc != ICP_B200_OK) return rc;
        ICP_CUDA(ctx, cudaEventRecord(ctx->ev[c][1], st));
        if (use_redo) {
            ICP_CUDA(ctx, cudaMemcpyAsync(ctx->h_redo + c, ctx->counters.as<uint32_t>() + kCtrRedoCount + c, 4, cudaMemcpyDeviceToHost, st));
            ICP_CUDA(ctx, cudaEventRecord(ctx->ev_cnt[c], st));
        }
        ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_T + (size_t)b0 * 6, ctx->out_T.as<double>() + (size_t)b0 * 6, (size_t)(b1 - b0) * 48, cudaMemcpyDeviceToHost, st));
        ctx->info.d2h_bytes += (uint64_t)(b1 - b0) * 48;
        if (hb->out_inner && hb->max_iter) {
            ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_inner + (size_t)b0 * hb->max_iter, ctx->out_inner.as<uint32_t>() + (size_t)b0 * hb->max_iter,
                                          (size_t)(b1 - b0) * hb->max_iter * 4, cudaMemcpyDeviceToHost, st));
            ctx->info.d2h_bytes += (uint64_t)(b1 - b0) * hb->max_iter * 4;
        }
        if (hb->out_corr_hash) {
            ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_corr_hash + b0, ctx->out_hash.as<uint64_t>() + b0, (size_t)(b1 - b0) * 8, cudaMemcpyDeviceToHost, st));
            ctx->info.d2h_bytes += (uint64_t)(b1 - b0) * 8;
        }
        if (hb->out_status) {
            ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_status + b0, ctx->out_status.as<int32_t>() + b0, (size_t)(b1 - b0) * 4, cudaMemcpyDeviceToHost, st));
            ctx->info.d2h_bytes += (uint64_t)(b1 - b0) * 4;
        }
        if (hb->out_cost) {
            ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_cost + b0, ctx->out_cost.as<double>() + b0, (size_t)(b1 - b0) * 8, cudaMemcpyDeviceToHost, st));
            ctx->info.d2h_bytes += (uint64_t)(b1 - b0) * 8;
        }
        if (ncorr) {
            const uint64_t mul = hb->corr_mode == ICP_B200_CORR_ALL ? hb->max_iter : 1;
            if (sp1 > sp0)
                ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_corr + sp0 * mul, ctx->out_corr.as<int32_t>() + sp0 * mul, (sp1 - sp0) * mul * 4, cudaMemcpyDeviceToHost, st));
            ctx->info.d2h_bytes += (sp1 - sp0) * mul * 4;
        }
        used = c + 1;
    }
    HostPending* P = new HostPending;
    P->hb = *hb; P->db = db; P->tgt = tgt; P->redo = redo; P->use_redo = use_redo; P->shared_src = shared_src;
    P->used = used; P->ncorr = ncorr; P->nsrc = nsrc; P->per = per;
    ctx->pending = P;
    return ICP_B200_OK;
}

// Second half: the guard's ordered re-run behind the last chunk, the status word, the results of re-run pairs.
int host_wait(icp_b200_context* ctx)
{
    if (!ctx->pending) return ICP_B200_OK;
    // the record is dropped however this ends: a failed wait leaves nothing pending
    struct Drop { icp_b200_context* c; ~Drop() { delete c->pending; c->pending = nullptr; } } drop{ctx};
    const HostPending& P = *ctx->pending;
    const icp_b200_batch* hb = &P.hb;
    const icp_b200_batch& db = P.db;
    icp_b200_target* tgt = P.tgt;
    const LaunchPlan& redo = P.redo;
    const bool use_redo = P.use_redo, shared_src = P.shared_src;
    const int used = P.used;
    const uint64_t ncorr = P.ncorr, nsrc = P.nsrc;
    const uint32_t B = hb->batch;
    const uint32_t* so = hb->src_off;
    cudaStream_t s0 = ctx->stream[0], s1 = ctx->stream[1];
    int rc = ICP_B200_OK;
    ICP_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t err = 0;
    uint32_t nredo = 0;
    std::vector<uint32_t> list;
    if (use_redo) {
        // Chunk by chunk, as the chunks end: a chunk that flagged pairs gets its ordered re-run at once, on the
        // re-run stream, beside the chunks still running (and beside the next batch of a pipelining caller), so the
        // 4-5 ms of a lone ordered pair are off the call's critical path unless the pair sits in the last chunk.
        bool any = false;
        std::vector<std::pair<uint32_t, uint32_t>> seg;    // (first pair of the chunk, flagged pairs)
        for (int c = 0; c < used; ++c) {
            ICP_CUDA(ctx, cudaEventSynchronize(ctx->ev_cnt[c]));
            const uint32_t n = ctx->h_redo[c];
            if (n == 0) continue;
            const uint32_t b0 = (uint32_t)c * P.per;
            if (!any) ICP_CUDA(ctx, cudaEventRecord(ctx->ev_redo[0], ctx->redo_stream));
            any = true;
            rc = launch_redo(ctx, db, tgt, redo, ctx->redo_stream, -1, false, c, b0, n);
            if (rc != ICP_B200_OK) return rc;
            seg.push_back({b0, n});
            nredo += n;
        }
        if (any) ICP_CUDA(ctx, cudaEventRecord(ctx->ev_redo[1], ctx->redo_stream));
        if ((rc = host_sync(ctx, s1)) != ICP_B200_OK) return rc;
        if ((rc = host_sync(ctx, s0)) != ICP_B200_OK) return rc;
        if (any) {
            list.resize(nredo);
            uint32_t at = 0;
            for (const auto& sg : seg) {
                ICP_CUDA(ctx, cudaMemcpyAsync(list.data() + at, ctx->redo_list.as<uint32_t>() + sg.first, (size_t)sg.second * 4,
                                              cudaMemcpyDeviceToHost, ctx->redo_stream));
                at += sg.second;
            }
            if ((rc = host_sync(ctx, ctx->redo_stream)) != ICP_B200_OK) return rc;
        }
    } else {
        if ((rc = host_sync(ctx, s1)) != ICP_B200_OK) return rc;
    }
    ICP_CUDA(ctx, cudaMemcpyAsync(&err, ctx->counters.as<int32_t>() + kCtrErr, 4, cudaMemcpyDeviceToHost, s0));
    if ((rc = host_sync(ctx, s0)) != ICP_B200_OK) return rc;
    ctx->info.redo_pairs = nredo;
    if (nredo > 0) {
        const bool all = nredo > 64;     // many (forced bands in tests): whole arrays again
        const size_t npairs = all ? 1 : nredo;
        for (size_t k = 0; k < npairs; ++k) {
            const uint32_t b0 = all ? 0u : list[k], cnt = all ? B : 1u;
            ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_T + (size_t)b0 * 6, ctx->out_T.as<double>() + (size_t)b0 * 6, (size_t)cnt * 48, cudaMemcpyDeviceToHost, s0));
            if (hb->out_inner && hb->max_iter)
                ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_inner + (size_t)b0 * hb->max_iter, ctx->out_inner.as<uint32_t>() + (size_t)b0 * hb->max_iter,
                                              (size_t)cnt * hb->max_iter * 4, cudaMemcpyDeviceToHost, s0));
            if (hb->out_corr_hash)
                ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_corr_hash + b0, ctx->out_hash.as<uint64_t>() + b0, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s0));
            if (hb->out_cost)
                ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_cost + b0, ctx->out_cost.as<double>() + b0, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s0));
            if (ncorr) {
                const uint64_t mul = hb->corr_mode == ICP_B200_CORR_ALL ? hb->max_iter : 1;
                const uint64_t sp0 = shared_src ? nsrc * b0 : so[b0], sp1 = shared_src ? nsrc * (b0 + cnt) : so[b0 + cnt];
                if (sp1 > sp0)
                    ICP_CUDA(ctx, cudaMemcpyAsync(hb->out_corr + sp0 * mul, ctx->out_corr.as<int32_t>() + sp0 * mul, (sp1 - sp0) * mul * 4, cudaMemcpyDeviceToHost, s0));
                ctx->info.d2h_bytes += (sp1 - sp0) * mul * 4;
            }
            ctx->info.d2h_bytes += (uint64_t)cnt * (48 + (hb->out_inner ? hb->max_iter * 4 : 0) + (hb->out_corr_hash ? 8 : 0) + (hb->out_cost ? 8 : 0));
        }
        if ((rc = host_sync(ctx, s0)) != ICP_B200_OK) return rc;
    }
    float total = 0.f;
    for (int c = 0; c < used; ++c) {
        float ms = 0.f;
        ICP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[c][0], ctx->ev[c][1]));
        total += ms;
    }
    if (nredo > 0) {
        float ms = 0.f;
        ICP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_redo[0], ctx->ev_redo[1]));
        total += ms;
    }
    ctx->info.kernel_ms = total;
    if (err != 0) return fail(ctx, pair_status_to_error(err), "at least one scan pair was rejected (see out_status)");
    return ICP_B200_OK;
}

int estimate_host_impl(icp_b200_context* ctx, icp_b200_target* tgt, const icp_b200_batch* hb)
{
    const int rc = host_submit(ctx, tgt, hb);
    return rc != ICP_B200_OK ? rc : host_wait(ctx);
}

}  // namespace

// =============================================================================================
extern "C" {

int icp_b200_abi_version(void) { return ICP_B200_ABI_VERSION; }

const char* icp_b200_status_string(int status)
{
    switch (status) {
        case ICP_B200_OK: return "ok";
        case ICP_B200_ERR_INVALID_ARGUMENT: return "invalid argument";
        case ICP_B200_ERR_CUDA: return "CUDA error";
        case ICP_B200_ERR_NO_DEVICE: return "no CUDA device";
        case ICP_B200_ERR_EMPTY_TARGET: return "empty target point set";
        case ICP_B200_ERR_NON_FINITE: return "non-finite input";
        case ICP_B200_ERR_TOO_LARGE: return "scan too large for this build";
        case ICP_B200_ERR_OUT_OF_MEMORY: return "out of device memory";
        default: return "unknown status";
    }
}

int icp_b200_device_count(int* count)
{
    if (!count) return ICP_B200_ERR_INVALID_ARGUMENT;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return ICP_B200_ERR_NO_DEVICE;
    }
    *count = n;
    return ICP_B200_OK;
}

int icp_b200_create(int device, icp_b200_context** out)
{
    if (!out) return ICP_B200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return ICP_B200_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) return ICP_B200_ERR_INVALID_ARGUMENT;
    icp_b200_context* ctx = new (std::nothrow) icp_b200_context();
    if (!ctx) return ICP_B200_ERR_OUT_OF_MEMORY;
    ctx->device = device;
    int prio_lo = 0, prio_hi = 0;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream[0], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream[1], cudaStreamNonBlocking) != cudaSuccess ||
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess ||
        // highest priority: a re-run CTA is placed ahead of the pending tree-mode CTAs when a slot frees up
        cudaStreamCreateWithPriority(&ctx->redo_stream, cudaStreamNonBlocking, env_int("ICP_B200_REDO_PRIORITY", 1) ? prio_hi : prio_lo) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return ICP_B200_ERR_CUDA;
    }
    ctx->cell_budget = std::max(16, std::min(65000, env_int("ICP_B200_CELL_BUDGET", 4096)));
    ctx->force_instance = env_int("ICP_B200_FORCE_INSTANCE", -1);
    ctx->stream_cells_per_point = std::max(0, std::min(256, env_int("ICP_B200_STREAM_CELLS_PER_POINT", 0)));
    ctx->force_stream = env_int("ICP_B200_FORCE_STREAM", 0);
    ctx->stream_nn_group = env_int("ICP_B200_STREAM_NN_GROUP", 0);
    ctx->stream_nn_f32 = env_int("ICP_B200_STREAM_NN_F32", 1);
    ctx->single_ordered = env_int("ICP_B200_SINGLE_ORDERED", 0);
    ctx->use_lean = env_int("ICP_B200_LEAN", 1);
    ctx->guard = env_int("ICP_B200_GUARD", 1);
    ctx->pdl = env_int("ICP_B200_PDL", 1);
    ctx->blocking_sync = env_int("ICP_B200_BLOCKING_SYNC", 0);
    ctx->fixed_point = env_int("ICP_B200_FIXED_POINT", 1);
    if (const char* gs = getenv("ICP_B200_GUARD_SCALE")) ctx->guard_scale = std::max(0.0, atof(gs));
    *out = ctx;
    return ICP_B200_OK;
}

void icp_b200_destroy(icp_b200_context* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    delete ctx->pending;   // a batch submitted and never waited for: its copies have landed by now
    ctx->pending = nullptr;
    DevBuf* bufs[] = {&ctx->src, &ctx->dst, &ctx->src_off, &ctx->dst_off, &ctx->init_T, &ctx->out_T, &ctx->out_inner,
                      &ctx->out_corr, &ctx->out_hash, &ctx->out_status, &ctx->counters, &ctx->workspace[0],
                      &ctx->workspace[1], &ctx->stats_in, &ctx->stats_off, &ctx->stats_out, &ctx->out_cost,
                      &ctx->redo_list, &ctx->redo_workspace, &ctx->dev_counters, &ctx->dev_workspace[0],
                      &ctx->dev_workspace[1], &ctx->dev_redo_list[0], &ctx->dev_redo_list[1],
                      &ctx->dev_redo_workspace[0], &ctx->dev_redo_workspace[1]};
    for (DevBuf* b : bufs) b->release();
    if (ctx->events_ok) {
        for (int i = 0; i < kMaxChunks; ++i)
            for (int j = 0; j < 2; ++j) cudaEventDestroy(ctx->ev[i][j]);
        cudaEventDestroy(ctx->ev_ready);
        cudaEventDestroy(ctx->ev_block);
        for (int i = 0; i < kMaxChunks; ++i) cudaEventDestroy(ctx->ev_cnt[i]);
        cudaFreeHost(ctx->h_redo);
        for (int j = 0; j < 2; ++j) cudaEventDestroy(ctx->ev_redo[j]);
    }
    for (int s = 0; s < 2; ++s)
        if (ctx->stream[s]) cudaStreamDestroy(ctx->stream[s]);
    if (ctx->redo_stream) cudaStreamDestroy(ctx->redo_stream);
    delete ctx;
}

int icp_b200_last_call_info(const icp_b200_context* ctx, icp_b200_call_info* info)
{
    if (!ctx || !info) return ICP_B200_ERR_INVALID_ARGUMENT;
    *info = ctx->info;
    return ICP_B200_OK;
}

const char* icp_b200_last_error(const icp_b200_context* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int icp_b200_estimate_batch(icp_b200_context* ctx, const icp_b200_batch* batch) { return estimate_host_impl(ctx, nullptr, batch); }

int icp_b200_estimate_batch_submit(icp_b200_context* ctx, const icp_b200_batch* batch) { return host_submit(ctx, nullptr, batch); }

int icp_b200_estimate_batch_wait(icp_b200_context* ctx)
{
    if (!ctx) return ICP_B200_ERR_INVALID_ARGUMENT;
    return host_wait(ctx);
}

int icp_b200_estimate_batch_device(icp_b200_context* ctx, const icp_b200_batch* batch, void* stream, int synchronize)
{
    return estimate_device_impl(ctx, nullptr, batch, stream, synchronize);
}

static int single_pair(icp_b200_context* ctx, uint32_t dim, const double* dst, uint32_t m, const double* src, uint32_t n,
                       const double* T0, uint32_t max_iter, double* out)
{
    if (!ctx || !T0 || !out) return ICP_B200_ERR_INVALID_ARGUMENT;
    const uint32_t so[2] = {0, n}, dof[2] = {0, m};
    icp_b200_batch b{};
    b.dim = dim; b.batch = 1; b.max_iter = max_iter; b.corr_mode = ICP_B200_CORR_NONE;
    b.src = src; b.src_off = so; b.dst = dst; b.dst_off = dof; b.init_T = T0; b.out_T = out;
    // One pair is a latency problem: the guarded warp-tree mode (same decisions, transform within 1e-9) finishes a
    // 666-point pair in well under half the time of the reference-order sums, whose 2n dependent additions per
    // evaluation cannot be spread over threads.  ICP_B200_SINGLE_ORDERED=1 asks for the bit-identical transform.
    b.flags = ctx->single_ordered ? ICP_B200_FLAG_ORDERED_SUMS : 0u;
    return estimate_host_impl(ctx, nullptr, &b);
}

int icp_b200_icp2d_estimate(icp_b200_context* ctx, const double* dst_xy, uint32_t n_dst, const double* src_xy,
                            uint32_t n_src, const double initial_transform[6], uint32_t max_iter, double out_transform[6])
{
    return single_pair(ctx, 2, dst_xy, n_dst, src_xy, n_src, initial_transform, max_iter, out_transform);
}

int icp_b200_icp3d_estimate(icp_b200_context* ctx, const double* dst_xyz, uint32_t n_dst, const double* src_xyz,
                            uint32_t n_src, const double initial_transform[6], uint32_t max_iter, double out_transform[6])
{
    return single_pair(ctx, 3, dst_xyz, n_dst, src_xyz, n_src, initial_transform, max_iter, out_transform);
}

// ---------------------------------------------------------------------------------------------
// (Re)builds the whole device structure of a target from t->host_pts: bounding box -> geometry -> counting sort ->
// binary32 copies -> jump-flooded seeds.  Used by icp_b200_target_create and, when an update leaves the
// geometry it was built for, by icp_b200_target_update.
static int target_build(icp_b200_target* t)
{
    icp_b200_context* ctx = t->ctx;
    const uint32_t dim = t->dim, m = (uint32_t)(t->host_pts.size() / dim);
    const double* dst = t->host_pts.data();
    t->m = m;
    t->seed_ptr = nullptr;
    cudaStream_t st = ctx->stream[0];
#define ICP_T(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cudaGetLastError(); return fail(ctx, e__ == cudaErrorMemoryAllocation ? ICP_B200_ERR_OUT_OF_MEMORY : ICP_B200_ERR_CUDA, #call, e__); } } while (0)
    if (m == 0) {  // an empty target is representable; estimating against it reports EMPTY_TARGET
        t->g = grid_make_geom(0, 0, 0, 0, 0, 0, (int)dim, 1);
        ICP_T(t->cstart.ensure(8));
        ICP_T(cudaMemsetAsync(t->cstart.p, 0, 8, st));
        ICP_T(cudaStreamSynchronize(st));
        return ICP_B200_OK;
    }
    ICP_T(t->pts.ensure((size_t)m * dim * 8));
    ICP_T(t->x.ensure((size_t)m * 8));
    ICP_T(t->y.ensure((size_t)m * 8));
    if (dim == 3) ICP_T(t->z.ensure((size_t)m * 8));
    ICP_T(t->idx.ensure((size_t)m * 4));
    ICP_T(t->slot.ensure((size_t)m * 4));
    ICP_T(t->keys.ensure(48));
    ICP_T(cudaMemcpyAsync(t->pts.p, dst, (size_t)m * dim * 8, cudaMemcpyHostToDevice, st));
    ctx->info.h2d_bytes += (uint64_t)m * dim * 8;
    unsigned long long keys[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
    ICP_T(cudaMemcpyAsync(t->keys.p, keys, 48, cudaMemcpyHostToDevice, st));
    TargetBuild tb{};
    tb.pts = t->pts.as<double>(); tb.m = m; tb.bbox_keys = t->keys.as<unsigned long long>();
    const int blocks = (int)std::min<uint32_t>((m + 255) / 256, (uint32_t)ctx->sm_count * 8);
    if (dim == 2) target_bbox_kernel<2><<<blocks, 256, 0, st>>>(tb); else target_bbox_kernel<3><<<blocks, 256, 0, st>>>(tb);
    ICP_T(cudaGetLastError());
    ICP_T(cudaMemcpyAsync(keys, t->keys.p, 48, cudaMemcpyDeviceToHost, st));
    ICP_T(cudaStreamSynchronize(st));
    auto unkey = [](unsigned long long k) {
        unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        double d; std::memcpy(&d, &b, 8); return d;
    };
    for (int q = 0; q < 6; ++q) t->bbox[q] = unkey(keys[q]);
    // cell budget per point, bounded (cell tables stay L2-resident).  2-D targets search a seeded disc
    // (icp_target_nn.cuh): empty cells cost nothing there, so the cells are small; the ring search of
    // 3-D targets pays for every empty cell it crosses.
    // 2-D: swept with the round-2 kernel on the 100k-point map (16.8 M scan queries / jittered map points, G queries/s):
    // 8: 6.66 / 9.00, 16: 6.61 / 9.75, 20: 6.74 / 10.27, 24: 6.60-6.67 / 10.5, 28: 6.06 / 9.89, 32: 6.30-6.36 / 10.1, 48: 5.49 / 9.58, 64: 4.87 / 9.32
    const int cells_per_point = env_int("ICP_B200_TARGET_CELLS_PER_POINT", dim == 2 ? 24 : 2);
    long long budget = std::max<long long>(64, std::min<long long>((long long)m * cells_per_point, 1ll << 24));
    t->g = grid_make_geom(t->bbox[0], t->bbox[1], t->bbox[2], t->bbox[3], t->bbox[4], t->bbox[5], (int)dim, (int)budget);
    t->m_built = m;
    const uint32_t ncells = (uint32_t)(t->g.gx * t->g.gy * t->g.gz);
    ICP_T(t->cnt.ensure((size_t)ncells * 4));
    ICP_T(t->cstart.ensure((size_t)(ncells + 1) * 4));
    ICP_T(cudaMemsetAsync(t->cnt.p, 0, (size_t)ncells * 4, st));
    tb.g = t->g; tb.cnt = t->cnt.as<uint32_t>(); tb.slot = t->slot.as<uint32_t>(); tb.cstart = t->cstart.as<uint32_t>();
    tb.x = t->x.as<double>(); tb.y = t->y.as<double>(); tb.z = t->z.as<double>(); tb.idx = t->idx.as<uint32_t>();
    // binary32 filter + seeds: 2-D, and a geometry whose squares stay inside binary32's range
    const bool fast2d = dim == 2 && env_int("ICP_B200_TARGET_FAST", 1) != 0 && m < kNnMaxPoints && t->g.h >= 1e-10 &&
                        t->g.h * (double)std::max(t->g.gx, t->g.gy) <= 1e15;
    if (fast2d) {
        ICP_T(t->pf.ensure((size_t)m * 8));
        ICP_T(t->seed.ensure((size_t)ncells * 4));
        tb.pf = t->pf.as<float2>();
    }
    if (dim == 2) target_count_kernel<2><<<blocks, 256, 0, st>>>(tb); else target_count_kernel<3><<<blocks, 256, 0, st>>>(tb);
    target_scan_kernel<<<1, 1024, 128, st>>>(tb, ncells);
    if (dim == 2) target_scatter_kernel<2><<<blocks, 256, 0, st>>>(tb); else target_scatter_kernel<3><<<blocks, 256, 0, st>>>(tb);
    ICP_T(cudaGetLastError());
    ctx->info.kernel_launches += 4;
    if (fast2d) {
        // jump flooding: steps 2^k >= half the grid ... 1, then 1 once more; cnt[] is free after the scan
        const int cblocks = (int)std::min<uint32_t>((ncells + 255) / 256, (uint32_t)ctx->sm_count * 16);
        uint32_t* bufs[2] = {t->seed.as<uint32_t>(), t->cnt.as<uint32_t>()};
        int cur = 0;
        target_seed_init_kernel<<<cblocks, 256, 0, st>>>(tb.cstart, ncells, bufs[cur]);
        int step = 1;
        while (step * 2 < std::max(t->g.gx, t->g.gy)) step *= 2;
        std::vector<int> steps;
        for (; step >= 1; step /= 2) steps.push_back(step);
        steps.push_back(1);
        ctx->info.kernel_launches += 1 + (uint32_t)steps.size();
        for (int sp : steps) {
            target_seed_pass_kernel<<<cblocks, 256, 0, st>>>(t->g, tb.x, tb.y, bufs[cur], bufs[cur ^ 1], sp);
            cur ^= 1;
        }
        if (cur != 0)
            ICP_T(cudaMemcpyAsync(bufs[0], bufs[1], (size_t)ncells * 4, cudaMemcpyDeviceToDevice, st));
        ICP_T(cudaGetLastError());
        t->seed_ptr = t->seed.as<uint32_t>();
    }
    ICP_T(cudaStreamSynchronize(st));
    return ICP_B200_OK;
}

int icp_b200_target_create(icp_b200_context* ctx, uint32_t dim, const double* dst, uint32_t m, icp_b200_target** out)
{
    if (!ctx || !out) return ICP_B200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (dim != 2 && dim != 3) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "dim must be 2 or 3");
    if (m && !dst) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null dst");
    ICP_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint64_t q = 0; q < (uint64_t)m * dim; ++q)
        if (!(std::fabs(dst[q]) <= kF64Max)) return fail(ctx, ICP_B200_ERR_NON_FINITE, "target holds NaN/Inf");
    icp_b200_target* t = new (std::nothrow) icp_b200_target();
    if (!t) return ICP_B200_ERR_OUT_OF_MEMORY;
    t->ctx = ctx; t->dim = dim;
    try {
        t->host_pts.assign(dst, dst + (size_t)m * dim);   // the library's copy of `dst` (the reference borrows it for 'a)
    } catch (const std::bad_alloc&) {
        delete t;
        return ICP_B200_ERR_OUT_OF_MEMORY;
    }
    ctx->info = icp_b200_call_info{};
    const int rc = target_build(t);
    if (rc != ICP_B200_OK) {
        icp_b200_target_destroy(t);
        return rc;
    }
    *out = t;
    return ICP_B200_OK;
}

// Icp2d::new on a CHANGED point set without paying for a new one (the reference rebuilds its kd-tree for
// every frame: src/lib.rs:97-102, examples/scan2d.rs:87).  New point set = the old one without the points
// `remove_idx` names (the survivors keep their order: indices close up), then `add` appended.
// As long as the added points lie inside the bounding box the grid was built for and the point count stays
// within a factor two of it, the update keeps the geometry and the cell tables' size: it re-sorts the points
// into the existing cells (count / scan / scatter), carries every cell's seed over to the new position of the
// point it named (a removed point hands over to its successor) and refloods the seeds once around the cells
// that received points — no bounding-box round trip, no allocation, 2 of the 14 seed passes.  Seeds only bound
// a search disc (icp_target_nn.cuh): whatever they are, searches return the exact nearest neighbour.
// Otherwise the structure is rebuilt (target_build).
int icp_b200_target_update(icp_b200_target* t, const double* add, uint32_t n_add, const uint32_t* remove_idx,
                           uint32_t n_remove)
{
    if (!t) return ICP_B200_ERR_INVALID_ARGUMENT;
    icp_b200_context* ctx = t->ctx;
    const uint32_t dim = t->dim, m_old = t->m;
    if ((n_add && !add) || (n_remove && !remove_idx)) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null add / remove_idx");
    for (uint64_t q = 0; q < (uint64_t)n_add * dim; ++q)
        if (!(std::fabs(add[q]) <= kF64Max)) return fail(ctx, ICP_B200_ERR_NON_FINITE, "added points hold NaN/Inf");
    ICP_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->info = icp_b200_call_info{};
    std::vector<uint8_t> gone(m_old, 0);
    for (uint32_t k = 0; k < n_remove; ++k) {
        if (remove_idx[k] >= m_old) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "remove_idx out of range");
        gone[remove_idx[k]] = 1;
    }
    // survivors close up; remap[old index] = new index of the point, for a removed point that of its successor
    std::vector<uint32_t> remap(m_old);
    std::vector<double> np;
    np.reserve(((size_t)m_old + n_add) * dim);
    uint32_t kept = 0;
    for (uint32_t j = 0; j < m_old; ++j) {
        remap[j] = kept;
        if (!gone[j]) {
            np.insert(np.end(), t->host_pts.begin() + (size_t)j * dim, t->host_pts.begin() + (size_t)(j + 1) * dim);
            ++kept;
        }
    }
    np.insert(np.end(), add, add + (size_t)n_add * dim);
    const uint32_t m = kept + n_add;
    bool inside = true;
    for (uint32_t k = 0; k < n_add && inside; ++k)
        for (uint32_t d = 0; d < dim; ++d)
            inside = inside && add[(size_t)k * dim + d] >= t->bbox[d] && add[(size_t)k * dim + d] <= t->bbox[3 + d];
    t->host_pts.swap(np);
    const bool incremental = inside && m_old > 0 && m > 0 && m >= t->m_built / 2 && m <= 2 * (uint64_t)t->m_built &&
                             env_int("ICP_B200_TARGET_INCREMENTAL", 1) != 0;
    if (!incremental) return target_build(t);
    for (uint32_t j = 0; j < m_old; ++j) remap[j] = std::min(remap[j], m - 1);   // a removed tail hands over to the last point

    cudaStream_t st = ctx->stream[0];
    const uint32_t ncells = (uint32_t)(t->g.gx * t->g.gy * t->g.gz);
    const bool fast2d = t->seed_ptr != nullptr;
    // the old order of the sorted points survives the re-sort in idx_prev (the seeds name sorted positions)
    if (fast2d) {
        ICP_CUDA(ctx, t->idx_prev.ensure((size_t)m_old * 4));
        ICP_CUDA(ctx, cudaMemcpyAsync(t->idx_prev.p, t->idx.p, (size_t)m_old * 4, cudaMemcpyDeviceToDevice, st));
        ICP_CUDA(ctx, t->remap.ensure((size_t)m_old * 4));
        ICP_CUDA(ctx, cudaMemcpyAsync(t->remap.p, remap.data(), (size_t)m_old * 4, cudaMemcpyHostToDevice, st));
        ICP_CUDA(ctx, t->where.ensure((size_t)m * 4));
    }
    ICP_CUDA(ctx, t->pts.ensure((size_t)m * dim * 8));
    ICP_CUDA(ctx, t->x.ensure((size_t)m * 8));
    ICP_CUDA(ctx, t->y.ensure((size_t)m * 8));
    if (dim == 3) ICP_CUDA(ctx, t->z.ensure((size_t)m * 8));
    ICP_CUDA(ctx, t->idx.ensure((size_t)m * 4));
    ICP_CUDA(ctx, t->slot.ensure((size_t)m * 4));
    if (fast2d) ICP_CUDA(ctx, t->pf.ensure((size_t)m * 8));
    ICP_CUDA(ctx, cudaMemcpyAsync(t->pts.p, t->host_pts.data(), (size_t)m * dim * 8, cudaMemcpyHostToDevice, st));
    ctx->info.h2d_bytes += (uint64_t)m * dim * 8 + (fast2d ? (uint64_t)m_old * 4 : 0);
    // seeds move to the second buffer while cnt[] (the flood's other buffer) counts the cells again
    if (fast2d) {
        ICP_CUDA(ctx, t->seed2.ensure((size_t)ncells * 4));
        ICP_CUDA(ctx, cudaMemcpyAsync(t->seed2.p, t->seed.p, (size_t)ncells * 4, cudaMemcpyDeviceToDevice, st));
    }
    ICP_CUDA(ctx, cudaMemsetAsync(t->cnt.p, 0, (size_t)ncells * 4, st));
    TargetBuild tb{};
    tb.pts = t->pts.as<double>(); tb.m = m; tb.g = t->g;
    tb.cnt = t->cnt.as<uint32_t>(); tb.slot = t->slot.as<uint32_t>(); tb.cstart = t->cstart.as<uint32_t>();
    tb.x = t->x.as<double>(); tb.y = t->y.as<double>(); tb.z = t->z.as<double>(); tb.idx = t->idx.as<uint32_t>();
    tb.pf = fast2d ? t->pf.as<float2>() : nullptr;
    tb.where = fast2d ? t->where.as<uint32_t>() : nullptr;
    const int blocks = (int)std::min<uint32_t>((m + 255) / 256, (uint32_t)ctx->sm_count * 8);
    if (dim == 2) target_count_kernel<2><<<blocks, 256, 0, st>>>(tb); else target_count_kernel<3><<<blocks, 256, 0, st>>>(tb);
    target_scan_kernel<<<1, 1024, 128, st>>>(tb, ncells);
    if (dim == 2) target_scatter_kernel<2><<<blocks, 256, 0, st>>>(tb); else target_scatter_kernel<3><<<blocks, 256, 0, st>>>(tb);
    ICP_CUDA(ctx, cudaGetLastError());
    ctx->info.kernel_launches += 3;
    if (fast2d) {
        const int cblocks = (int)std::min<uint32_t>((ncells + 255) / 256, (uint32_t)ctx->sm_count * 16);
        target_seed_remap_kernel<<<cblocks, 256, 0, st>>>(t->seed2.as<uint32_t>(), ncells, t->idx_prev.as<uint32_t>(), m_old,
                                                          t->remap.as<uint32_t>(), t->where.as<uint32_t>(), m, t->seed.as<uint32_t>());
        if (n_add)
            target_seed_added_kernel<<<(int)std::min<uint32_t>((n_add + 255) / 256, 4096u), 256, 0, st>>>(
                t->g, t->pts.as<double>(), kept, m, t->where.as<uint32_t>(), t->seed.as<uint32_t>());
        // one reflood around the touched cells: two passes of step 1 (seed -> seed2 -> seed)
        target_seed_pass_kernel<<<cblocks, 256, 0, st>>>(t->g, tb.x, tb.y, t->seed.as<uint32_t>(), t->seed2.as<uint32_t>(), 1);
        target_seed_pass_kernel<<<cblocks, 256, 0, st>>>(t->g, tb.x, tb.y, t->seed2.as<uint32_t>(), t->seed.as<uint32_t>(), 1);
        ICP_CUDA(ctx, cudaGetLastError());
        ctx->info.kernel_launches += 3 + (n_add ? 1 : 0);
    }
    ICP_CUDA(ctx, cudaStreamSynchronize(st));
    t->m = m;
    return ICP_B200_OK;
}

void icp_b200_target_destroy(icp_b200_target* t)
{
    if (!t) return;
    if (t->ctx) cudaSetDevice(t->ctx->device);
    DevBuf* bufs[] = {&t->pts, &t->x, &t->y, &t->z, &t->idx, &t->cstart, &t->cnt, &t->slot, &t->keys, &t->queries, &t->out_idx, &t->pf, &t->seed,
                      &t->idx_prev, &t->remap, &t->where, &t->seed2};
    for (DevBuf* b : bufs) b->release();
    delete t;
}

int icp_b200_target_estimate_batch(icp_b200_target* t, const icp_b200_batch* batch)
{
    if (!t) return ICP_B200_ERR_INVALID_ARGUMENT;
    if (batch && batch->dim != t->dim) return fail(t->ctx, ICP_B200_ERR_INVALID_ARGUMENT, "batch dim differs from target dim");
    return estimate_host_impl(t->ctx, t, batch);
}

int icp_b200_target_estimate_batch_submit(icp_b200_target* t, const icp_b200_batch* batch)
{
    if (!t) return ICP_B200_ERR_INVALID_ARGUMENT;
    if (batch && batch->dim != t->dim) return fail(t->ctx, ICP_B200_ERR_INVALID_ARGUMENT, "batch dim differs from target dim");
    return host_submit(t->ctx, t, batch);
}

int icp_b200_target_estimate_batch_device(icp_b200_target* t, const icp_b200_batch* batch, void* stream, int synchronize)
{
    if (!t) return ICP_B200_ERR_INVALID_ARGUMENT;
    if (batch && batch->dim != t->dim) return fail(t->ctx, ICP_B200_ERR_INVALID_ARGUMENT, "batch dim differs from target dim");
    return estimate_device_impl(t->ctx, t, batch, stream, synchronize);
}

int icp_b200_target_search_device(icp_b200_target* t, const double* queries, uint64_t nq, int32_t* out_index, void* stream_v,
                                  int synchronize)
{
    if (!t) return ICP_B200_ERR_INVALID_ARGUMENT;
    icp_b200_context* ctx = t->ctx;
    if (nq == 0) return ICP_B200_OK;
    if (!queries || !out_index) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null queries / out_index");
    if (t->m == 0) return fail(ctx, ICP_B200_ERR_EMPTY_TARGET, "search on an empty target");
    if (t->dim == 2 && (reinterpret_cast<uintptr_t>(queries) & 15u) != 0)
        return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "2-D queries must be 16-byte aligned (they are read as 128-bit loads)");
    ICP_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_events(ctx);
    if (rc != ICP_B200_OK) return rc;
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : ctx->stream[0];
    TargetView tv{};
    tv.x = t->x.as<double>(); tv.y = t->y.as<double>(); tv.z = t->z.as<double>();
    tv.idx = t->idx.as<uint32_t>(); tv.cstart = t->cstart.as<uint32_t>(); tv.g = t->g; tv.m = t->m;
    tv.seed = t->seed_ptr; tv.pf = t->seed_ptr ? t->pf.as<float2>() : nullptr;
    const int threads = 256;
    const long long want = (long long)((nq + threads - 1) / threads);
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)ctx->sm_count * 8));
    ICP_CUDA(ctx, cudaEventRecord(ctx->ev[0][0], st));
    // ICP_B200_NN_WARP=1: the warp-cooperative union walk (measured slower on scan queries: 3.55 against 5.90 G
    // queries/s — consecutive beams are ~60 mm apart at 10 m, their discs barely overlap; kept for A/B runs)
    const int nn_group = (t->dim == 2 && tv.pf != nullptr) ? env_int("ICP_B200_NN_GROUP", 0) : 0;
    if (nn_group == 4 || nn_group == 8) {
        // ICP_B200_NN_GROUP=4|8: G lanes per query (target_nn_search2_group), a warp takes 32 / G consecutive queries
        // at a time.  Measured on the config-3 scan queries: 4.08 (G = 4) and 3.31 (G = 8) against 6.41 G queries/s
        // for the private walk; kept for A/B runs.
        const long long chunks = (long long)((nq + (32 / nn_group) - 1) / (32 / nn_group));
        const long long wantg = (chunks + kNnGroupThreads / 32 - 1) / (kNnGroupThreads / 32);
        const int gridg = (int)std::max<long long>(1, std::min<long long>(wantg, (long long)ctx->sm_count * 32));
        if (nn_group == 4) nn_search_group_kernel<4><<<gridg, kNnGroupThreads, 0, st>>>(tv, queries, nq, out_index);
        else nn_search_group_kernel<8><<<gridg, kNnGroupThreads, 0, st>>>(tv, queries, nq, out_index);
    } else if (t->dim == 2 && tv.pf != nullptr && env_int("ICP_B200_NN_WARP", 0) != 0) {
        const long long wantw = (long long)((nq + kNnWarpThreads - 1) / kNnWarpThreads);
        const int gridw = (int)std::max<long long>(1, std::min<long long>(wantw, (long long)ctx->sm_count * 16));
        nn_search_warp_kernel<<<gridw, kNnWarpThreads, kNnWarpShared, st>>>(tv, queries, nq, out_index);
        ctx->info.grid_ctas = (uint32_t)gridw;
    } else if (t->dim == 2) nn_search_kernel<2><<<grid, threads, 0, st>>>(tv, queries, nq, out_index);
    else nn_search_kernel<3><<<grid, threads, 0, st>>>(tv, queries, nq, out_index);
    ICP_CUDA(ctx, cudaGetLastError());
    ICP_CUDA(ctx, cudaEventRecord(ctx->ev[0][1], st));
    ctx->info.kernel_launches += 1;
    if (!(t->dim == 2 && tv.pf != nullptr && env_int("ICP_B200_NN_WARP", 0) != 0)) ctx->info.grid_ctas = (uint32_t)grid;
    ctx->info.block_threads = threads; ctx->info.smem_bytes = 0;
    if (synchronize) {
        ICP_CUDA(ctx, cudaStreamSynchronize(st));
        ICP_CUDA(ctx, cudaEventElapsedTime(&ctx->info.kernel_ms, ctx->ev[0][0], ctx->ev[0][1]));
    }
    return ICP_B200_OK;
}

int icp_b200_target_search(icp_b200_target* t, const double* queries, uint64_t nq, int32_t* out_index)
{
    if (!t) return ICP_B200_ERR_INVALID_ARGUMENT;
    icp_b200_context* ctx = t->ctx;
    if (nq == 0) return ICP_B200_OK;
    if (!queries || !out_index) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null queries / out_index");
    ICP_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->info = icp_b200_call_info{};
    ICP_CUDA(ctx, t->queries.ensure(nq * t->dim * 8));
    ICP_CUDA(ctx, t->out_idx.ensure(nq * 4));
    cudaStream_t st = ctx->stream[0];
    ICP_CUDA(ctx, cudaMemcpyAsync(t->queries.p, queries, nq * t->dim * 8, cudaMemcpyHostToDevice, st));
    int rc = icp_b200_target_search_device(t, t->queries.as<double>(), nq, t->out_idx.as<int32_t>(), st, 0);
    if (rc != ICP_B200_OK) return rc;
    ICP_CUDA(ctx, cudaMemcpyAsync(out_index, t->out_idx.p, nq * 4, cudaMemcpyDeviceToHost, st));
    ICP_CUDA(ctx, cudaStreamSynchronize(st));
    ICP_CUDA(ctx, cudaEventElapsedTime(&ctx->info.kernel_ms, ctx->ev[0][0], ctx->ev[0][1]));
    ctx->info.h2d_bytes = nq * t->dim * 8;
    ctx->info.d2h_bytes = nq * 4;
    return ICP_B200_OK;
}

// ---------------------------------------------------------------------------------------------
int icp_b200_calc_stddevs_batch(icp_b200_context* ctx, const double* residuals_xy, const uint32_t* off, uint32_t batch,
                                double* out_sigma, double* out_median)
{
    if (!ctx) return ICP_B200_ERR_INVALID_ARGUMENT;
    if (batch == 0) return ICP_B200_OK;
    if (!off || !out_sigma) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null off / out_sigma");
    ICP_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->info = icp_b200_call_info{};
    int cap = 1;
    for (uint32_t b = 0; b < batch; ++b) cap = std::max(cap, (int)(off[b + 1] - off[b]));
    const uint32_t smem = stats_smem_bytes(cap);
    if ((int)smem > ctx->max_smem_optin || cap > kStatsThreads * kStatsPpt)
        return fail(ctx, ICP_B200_ERR_TOO_LARGE, "residual set larger than 3072 points");
    const uint64_t total = off[batch];
    if (total && !residuals_xy) return fail(ctx, ICP_B200_ERR_INVALID_ARGUMENT, "null residuals");
    ICP_CUDA(ctx, ctx->stats_in.ensure(std::max<uint64_t>(total, 1) * 16));
    ICP_CUDA(ctx, ctx->stats_off.ensure((batch + 1) * 4));
    ICP_CUDA(ctx, ctx->stats_out.ensure((size_t)batch * 32));
    cudaStream_t st = ctx->stream[0];
    if (total) ICP_CUDA(ctx, cudaMemcpyAsync(ctx->stats_in.p, residuals_xy, total * 16, cudaMemcpyHostToDevice, st));
    ICP_CUDA(ctx, cudaMemcpyAsync(ctx->stats_off.p, off, (batch + 1) * 4, cudaMemcpyHostToDevice, st));
    StatsParams p{};
    p.residuals = ctx->stats_in.as<double>(); p.off = ctx->stats_off.as<uint32_t>();
    p.out_sigma = ctx->stats_out.as<double>(); p.out_median = ctx->stats_out.as<double>() + (size_t)batch * 2;
    p.B = batch; p.cap = cap; p.mode = env_int("ICP_B200_STATS_MODE", 1);
    ICP_CUDA(ctx, cudaFuncSetAttribute(calc_stddevs_kernel<kStatsThreads, kStatsPpt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<uint32_t>(batch, (uint32_t)ctx->sm_count * 4);
    calc_stddevs_kernel<kStatsThreads, kStatsPpt><<<grid, kStatsThreads, smem, st>>>(p);
    ICP_CUDA(ctx, cudaGetLastError());
    ICP_CUDA(ctx, cudaMemcpyAsync(out_sigma, p.out_sigma, (size_t)batch * 16, cudaMemcpyDeviceToHost, st));
    if (out_median) ICP_CUDA(ctx, cudaMemcpyAsync(out_median, p.out_median, (size_t)batch * 16, cudaMemcpyDeviceToHost, st));
    ICP_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->info.kernel_launches = 1;
    return ICP_B200_OK;
}

}  // extern "C"
