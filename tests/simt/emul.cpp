// emul.cpp — runs the product's kernel headers on the lock-step CTA emulator (cpu_simt.h).
// TEST INFRASTRUCTURE: built into tests/simt/_build/libicp_simt.so by tests/simt/Makefile and
// loaded only by the CPU-only tests.  Never shipped, never a fallback.
#define ICP_SIMT_EMULATE 1
#include "icp_lean.cuh"
#include "icp_stream.cuh"
#include "icp_stats_kernel.cuh"
#include "icp_target.cuh"

#include <vector>

using namespace icp;

// ordered: 0 warp-tree (icp_fused.cuh), 1 ordered sums, 2 warp-tree on the icp_lean.cuh schedule
template <int THREADS, int PPT, int DIM>
static void run_fused(FusedParams p, unsigned grid, int ordered)
{
    if (ordered == 2 || ordered == 3) {   // lean schedule; 3: dedicated service warps (capacity (THREADS-64)*PPT)
        if constexpr (THREADS >= 128) {
            if (ordered == 3) {
                const LeanLayout LL = lean_layout(DIM, (THREADS - 64) * PPT, p.mcap, p.cell_budget, false);
                simt::launch(grid, THREADS, LL.total, [&] { icp_lean_kernel<THREADS, PPT, DIM, false, 1, true>(p); });
                return;
            }
        }
        if constexpr (THREADS >= 64) {
            const LeanLayout LL = lean_layout(DIM, THREADS * PPT, p.mcap, p.cell_budget, false);
            simt::launch(grid, THREADS, LL.total, [&] { icp_lean_kernel<THREADS, PPT, DIM, false, 1, false>(p); });
        }
        return;
    }
    const FusedLayout L = fused_layout(DIM, THREADS * PPT, p.mcap, p.cell_budget, false, ordered);
    if (ordered) simt::launch(grid, THREADS, L.total, [&] { icp_fused_kernel<THREADS, PPT, DIM, false, 1, true>(p); });
    else simt::launch(grid, THREADS, L.total, [&] { icp_fused_kernel<THREADS, PPT, DIM, false, 1, false>(p); });
}

extern "C" int simt_set_schedule(int mode)
{
    simt::st().schedule = mode;
    return 0;
}

// Options of the next simt_icp_fused / simt_icp_stream calls (the host library's per-call state):
// guard_scale > 0: tree-mode runs use the knife-edge guard with the library's bands times that
// scale and re-run flagged pairs with ordered sums, exactly as icp_b200.cu's launch_range does;
// out_cost: where the final huber_error goes; stream_ordered: reference-order sums in the
// streaming kernel; shared_src: ICP_B200_FLAG_SHARED_SOURCE.
static double g_guard_scale = 0.0, g_band_delta_scale = 1.0;
static double* g_out_cost = nullptr;
static int g_stream_ordered = 0, g_shared_src = 0, g_stream_nn_group = 0, g_fixed_point = 1;
extern "C" void simt_set_fixed_point(int on) { g_fixed_point = on; }
extern "C" void simt_set_stream_nn_group(int v) { g_stream_nn_group = v; }
static int g_stream_nn_f32 = 1;
extern "C" void simt_set_stream_nn_f32(int v) { g_stream_nn_f32 = v; }
static uint32_t g_last_redo = 0;
extern "C" void simt_set_options(double guard_scale, double* out_cost, int stream_ordered, int shared_src,
                                 double delta_band_scale)
{
    g_guard_scale = guard_scale; g_out_cost = out_cost; g_stream_ordered = stream_ordered; g_shared_src = shared_src;
    g_band_delta_scale = delta_band_scale;   // extra factor on the delta band only (tests isolate the error band with it)
}
extern "C" uint32_t simt_last_redo_count() { return g_last_redo; }
static void apply_guard(FusedParams& p, std::vector<uint32_t>& list, uint32_t* count, uint32_t B)
{
    list.assign(B + 1, 0u);
    *count = 0;
    p.redo_list = list.data(); p.redo_count = count;
    p.guard.det = kGuardDet * g_guard_scale; p.guard.delta = kGuardDelta * g_guard_scale * g_band_delta_scale; p.guard.err = kGuardErr * g_guard_scale;
}

// Same argument meaning as icp_b200_batch.  variant: 0 -> 128 threads x 3 points, 1 -> 128 x 9,
// 2 -> 64 x 2 (tiny), 3 -> 32 x 4 (single warp), 4 -> 256 x 6.
extern "C" int simt_icp_fused(int dim, int variant, unsigned grid, const double* src, const uint32_t* src_off,
                              const double* dst, const uint32_t* dst_off, const double* init_T, uint32_t B,
                              uint32_t max_iter, int corr_mode, int cell_budget, int ordered, double* out_T, uint32_t* out_inner,
                              int32_t* out_corr, uint64_t* out_hash, int32_t* out_status)
{
    uint32_t counter = 0;
    int32_t err = 0;
    int mcap = 0;
    for (uint32_t b = 0; b < B; ++b) mcap = std::max(mcap, (int)(dst_off[b + 1] - dst_off[b]));
    mcap = (mcap + 7) & ~7;
    FusedParams p{};
    p.src = src; p.src_off = src_off; p.dst = dst; p.dst_off = dst_off; p.init_T = init_T;
    p.out_T = out_T; p.out_inner = out_inner; p.out_corr = out_corr; p.out_hash = out_hash;
    p.out_status = out_status; p.out_error = &err; p.work_counter = &counter;
    p.pair_begin = 0; p.pair_end = B; p.max_iter = max_iter; p.corr_mode = corr_mode;
    p.cell_budget = cell_budget; p.mcap = mcap; p.sm_count = 1;
    p.out_cost = g_out_cost; p.shared_src = (uint32_t)g_shared_src; p.fixed_point = (uint32_t)g_fixed_point;
    const FusedLayout L = fused_layout(dim, 256 * 6, mcap, cell_budget, false, true);
    std::vector<unsigned char> ws((size_t)grid * (L.nn_bytes + 64));
    p.workspace = ws.data();
    p.ws_stride = L.nn_bytes + 64 - ((L.nn_bytes + 64) % 16);
    std::vector<uint32_t> redo_list;
    uint32_t redo_count = 0, counter2 = 0;
    const bool guard = g_guard_scale > 0.0 && ordered != 1;
    if (guard) apply_guard(p, redo_list, &redo_count, B);
    auto run = [&](const FusedParams& q, int mode) {
        FusedParams r = q;
        if (dim == 2) {
            if (variant == 0) { r.ncap = 384; run_fused<128, 3, 2>(r, grid, mode); }
            else if (variant == 1) { r.ncap = 1152; run_fused<128, 9, 2>(r, grid, mode); }
            else if (variant == 2) { r.ncap = 128; run_fused<64, 2, 2>(r, grid, mode); }
            else if (variant == 4) { r.ncap = 1536; run_fused<256, 6, 2>(r, grid, mode); }
            else { r.ncap = 128; run_fused<32, 4, 2>(r, grid, mode); }
        } else {
            if (variant == 0) { r.ncap = 384; run_fused<128, 3, 3>(r, grid, mode); }
            else if (variant == 1) { r.ncap = 1152; run_fused<128, 9, 3>(r, grid, mode); }
            else if (variant == 2) { r.ncap = 128; run_fused<64, 2, 3>(r, grid, mode); }
            else if (variant == 4) { r.ncap = 1536; run_fused<256, 6, 3>(r, grid, mode); }
            else { r.ncap = 128; run_fused<32, 4, 3>(r, grid, mode); }
        }
    };
    // self-cleaning counters (device-pointer calls of the library): the last CTA zeroes them
    uint32_t fin = 0, fin2 = 0, seen = 0;
    p.fin_counter = &fin;
    run(p, ordered);
    if (counter != 0 || fin != 0) return -98;
    g_last_redo = 0;
    if (guard) {   // the ordered re-run of the flagged pairs (icp_b200.cu launch_range)
        FusedParams r = p;
        r.pair_list = redo_list.data(); r.pair_count = &redo_count;
        r.redo_list = nullptr; r.redo_count = nullptr; r.guard = GuardBands{0., 0., 0.};
        r.work_counter = &counter2;
        r.fin_counter = &fin2; r.redo_seen = &seen;
        run(r, 1);
        if (counter2 != 0 || fin2 != 0 || redo_count != 0) return -98;
        g_last_redo = seen;
    }
    return err;
}

// Streaming (large-scan) kernel on the emulator; threads = 32, 64 or 128.
template <int THREADS, int DIM>
static void run_stream(StreamParams sp, unsigned grid)
{
    simt::launch(grid, THREADS, stream_smem_bytes(sp.ordered != 0), [&] { icp_stream_kernel<THREADS, DIM, false>(sp); });
}

extern "C" int simt_icp_stream(int dim, int threads, unsigned grid, const double* src, const uint32_t* src_off,
                               const double* dst, const uint32_t* dst_off, const double* init_T, uint32_t B,
                               uint32_t max_iter, int corr_mode, double* out_T, uint32_t* out_inner, int32_t* out_corr,
                               uint64_t* out_hash, int32_t* out_status)
{
    uint32_t counter = 0;
    int32_t err = 0;
    uint32_t ncap = 8, mcap = 8;
    for (uint32_t b = 0; b < B; ++b) {
        ncap = std::max(ncap, src_off[b + 1] - src_off[b]);
        mcap = std::max(mcap, dst_off[b + 1] - dst_off[b]);
    }
    ncap = (ncap + 7u) & ~7u;
    mcap = (mcap + 7u) & ~7u;
    StreamParams sp{};
    FusedParams& p = sp.f;
    p.src = src; p.src_off = src_off; p.dst = dst; p.dst_off = dst_off; p.init_T = init_T;
    p.out_T = out_T; p.out_inner = out_inner; p.out_corr = out_corr; p.out_hash = out_hash;
    p.out_status = out_status; p.out_error = &err; p.work_counter = &counter;
    p.pair_begin = 0; p.pair_end = B; p.max_iter = max_iter; p.corr_mode = corr_mode; p.sm_count = 1;
    sp.n_cap = ncap; sp.m_cap = mcap; sp.c_cap = std::max(64u, 2u * mcap);
    p.out_cost = g_out_cost; p.shared_src = (uint32_t)g_shared_src; p.fixed_point = (uint32_t)g_fixed_point;
    const StreamLayout L = stream_layout(dim, sp.n_cap, sp.m_cap, sp.c_cap, false, true);
    std::vector<unsigned char> ws((size_t)grid * L.total + 256, 0xCD);
    p.workspace = ws.data();
    p.ws_stride = L.total;
    std::vector<uint32_t> redo_list;
    uint32_t redo_count = 0, counter2 = 0;
    const bool guard = g_guard_scale > 0.0 && !g_stream_ordered;
    if (guard) apply_guard(p, redo_list, &redo_count, B);
    auto run = [&](StreamParams q) {
        if (dim == 2) {
            if (threads == 32) run_stream<32, 2>(q, grid);
            else if (threads == 64) run_stream<64, 2>(q, grid);
            else run_stream<128, 2>(q, grid);
        } else {
            if (threads == 32) run_stream<32, 3>(q, grid);
            else if (threads == 64) run_stream<64, 3>(q, grid);
            else run_stream<128, 3>(q, grid);
        }
    };
    sp.ordered = (uint32_t)g_stream_ordered;
    sp.nn_group = (uint32_t)g_stream_nn_group;
    sp.nn_f32 = (uint32_t)g_stream_nn_f32;
    run(sp);
    g_last_redo = 0;
    if (guard) {
        StreamParams r = sp;
        r.ordered = 1u;
        r.f.pair_list = redo_list.data(); r.f.pair_count = &redo_count;
        r.f.redo_list = nullptr; r.f.redo_count = nullptr; r.f.guard = GuardBands{0., 0., 0.};
        r.f.work_counter = &counter2;
        run(r);
        g_last_redo = redo_count;
    }
    return err;
}

extern "C" int simt_calc_stddevs(const double* residuals, const uint32_t* off, uint32_t B, int mode,
                                 int threads, double* out_sigma, double* out_median)
{
    int cap = 1;
    for (uint32_t b = 0; b < B; ++b) cap = std::max(cap, (int)(off[b + 1] - off[b]));
    StatsParams p{};
    p.residuals = residuals; p.off = off; p.out_sigma = out_sigma; p.out_median = out_median;
    p.B = B; p.cap = cap; p.mode = mode;
    const unsigned grid = B < 4 ? (B ? B : 1) : 4;
    if (threads == 32) simt::launch(grid, 32, stats_smem_bytes(cap), [&] { calc_stddevs_kernel<32, 40>(p); });
    else if (threads == 256) simt::launch(grid, 256, stats_smem_bytes(cap), [&] { calc_stddevs_kernel<256, 8>(p); });
    else simt::launch(grid, 128, stats_smem_bytes(cap), [&] { calc_stddevs_kernel<128, kStatsPpt>(p); });
    return 0;
}

// Shared-target structure (icp_target.cuh) built and searched on the emulator: the same kernels the
// library launches in icp_b200_target_create / icp_b200_target_search.  fast != 0 selects the seeded
// binary32-filtered search of icp_target_nn.cuh (2-D; fast == 2: its warp-cooperative kernel), else the ring search.
extern "C" int simt_target_search(int dim, const double* pts, uint32_t m, long long budget, int fast,
                                  const double* queries, uint64_t nq, int32_t* out_idx)
{
    if (m == 0 || (dim != 2 && dim != 3)) return -1;
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) {
        mn[d] = mx[d] = pts[d];
        for (uint32_t j = 1; j < m; ++j) {
            mn[d] = std::min(mn[d], pts[(size_t)j * dim + d]);
            mx[d] = std::max(mx[d], pts[(size_t)j * dim + d]);
        }
    }
    const GridGeom g = grid_make_geom(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], dim, (int)budget);
    const uint32_t ncells = (uint32_t)(g.gx * g.gy * g.gz);
    std::vector<uint32_t> cnt(ncells, 0u), slot(m), cstart(ncells + 1), idx(m), seed(ncells), seed2(ncells);
    std::vector<double> x(m), y(m), z(m);
    std::vector<float2> pf(m);
    TargetBuild tb{};
    tb.pts = pts; tb.m = m; tb.g = g; tb.cnt = cnt.data(); tb.slot = slot.data(); tb.cstart = cstart.data();
    tb.x = x.data(); tb.y = y.data(); tb.z = z.data(); tb.idx = idx.data();
    const bool fast2d = fast && dim == 2;
    tb.pf = fast2d ? pf.data() : nullptr;
    if (dim == 2) simt::launch(3, 256, 0, [&] { target_count_kernel<2>(tb); });
    else simt::launch(3, 256, 0, [&] { target_count_kernel<3>(tb); });
    simt::launch(1, 1024, 128, [&] { target_scan_kernel(tb, ncells); });
    if (dim == 2) simt::launch(3, 256, 0, [&] { target_scatter_kernel<2>(tb); });
    else simt::launch(3, 256, 0, [&] { target_scatter_kernel<3>(tb); });
    TargetView tv{};
    tv.x = x.data(); tv.y = y.data(); tv.z = z.data(); tv.idx = idx.data(); tv.cstart = cstart.data();
    tv.g = g; tv.m = m;
    if (fast2d) {
        uint32_t* bufs[2] = {seed.data(), seed2.data()};
        int cur = 0;
        simt::launch(4, 256, 0, [&] { target_seed_init_kernel(cstart.data(), ncells, bufs[0]); });
        int step = 1;
        while (step * 2 < std::max(g.gx, g.gy)) step *= 2;
        std::vector<int> steps;
        for (; step >= 1; step /= 2) steps.push_back(step);
        steps.push_back(1);
        for (int sp : steps) {
            simt::launch(4, 256, 0, [&] { target_seed_pass_kernel(g, x.data(), y.data(), bufs[cur], bufs[cur ^ 1], sp); });
            cur ^= 1;
        }
        tv.seed = bufs[cur]; tv.pf = pf.data();
    }
    if (dim == 2 && fast == 2) simt::launch(2, kNnWarpThreads, kNnWarpShared, [&] { nn_search_warp_kernel(tv, queries, nq, out_idx); });
    else if (dim == 2 && fast == 4) simt::launch(2, kNnGroupThreads, 0, [&] { nn_search_group_kernel<4>(tv, queries, nq, out_idx); });
    else if (dim == 2 && fast == 8) simt::launch(2, kNnGroupThreads, 0, [&] { nn_search_group_kernel<8>(tv, queries, nq, out_idx); });
    else if (dim == 2) simt::launch(2, 256, 0, [&] { nn_search_kernel<2>(tv, queries, nq, out_idx); });
    else simt::launch(2, 256, 0, [&] { nn_search_kernel<3>(tv, queries, nq, out_idx); });
    return 0;
}

// Seeded search of icp_target_nn.cuh (grid_nn_search_auto) on a sorted cell list built by the target
// kernels, one emulated thread per query.  seeds[q]: ORIGINAL index of any target point, or -1
// (own-cell seed / ring search).  2-D and 3-D.
template <int DIM>
static void seeded_kernel(GridView<DIM, uint32_t, uint32_t> gv, const double* queries, const int32_t* seed_pos,
                          unsigned long long nq, int32_t* out_idx)
{
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq;
         q += (unsigned long long)gridDim.x * blockDim.x) {
        const double qz = DIM == 3 ? queries[q * DIM + 2] : 0.0;
        const NnResult<DIM, uint32_t, uint32_t> r = grid_nn_search_auto(gv, queries[q * DIM], queries[q * DIM + 1], qz, seed_pos[q]);
        out_idx[q] = r.idx;
    }
}

extern "C" int simt_grid_seeded_search(int dim, const double* pts, uint32_t m, long long budget, const double* queries,
                                       const int32_t* seeds, uint64_t nq, int32_t* out_idx)
{
    if (m == 0 || (dim != 2 && dim != 3)) return -1;
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) {
        mn[d] = mx[d] = pts[d];
        for (uint32_t j = 1; j < m; ++j) {
            mn[d] = std::min(mn[d], pts[(size_t)j * dim + d]);
            mx[d] = std::max(mx[d], pts[(size_t)j * dim + d]);
        }
    }
    const GridGeom g = grid_make_geom(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], dim, (int)budget);
    const uint32_t ncells = (uint32_t)(g.gx * g.gy * g.gz);
    std::vector<uint32_t> cnt(ncells, 0u), slot(m), cstart(ncells + 1), idx(m), where(m);
    std::vector<double> x(m), y(m), z(m);
    TargetBuild tb{};
    tb.pts = pts; tb.m = m; tb.g = g; tb.cnt = cnt.data(); tb.slot = slot.data(); tb.cstart = cstart.data();
    tb.x = x.data(); tb.y = y.data(); tb.z = z.data(); tb.idx = idx.data(); tb.pf = nullptr;
    if (dim == 2) simt::launch(2, 256, 0, [&] { target_count_kernel<2>(tb); });
    else simt::launch(2, 256, 0, [&] { target_count_kernel<3>(tb); });
    simt::launch(1, 1024, 128, [&] { target_scan_kernel(tb, ncells); });
    if (dim == 2) simt::launch(2, 256, 0, [&] { target_scatter_kernel<2>(tb); });
    else simt::launch(2, 256, 0, [&] { target_scatter_kernel<3>(tb); });
    for (uint32_t p = 0; p < m; ++p) where[idx[p]] = p;   // original index -> sorted position
    std::vector<int32_t> seed_pos(nq);
    for (uint64_t q = 0; q < nq; ++q) seed_pos[q] = seeds[q] < 0 ? -1 : (int32_t)where[(uint32_t)seeds[q] % m];
    if (dim == 2) {
        GridView<2, uint32_t, uint32_t> gv;
        gv.x = x.data(); gv.y = y.data(); gv.z = z.data(); gv.idx = idx.data(); gv.cstart = cstart.data(); gv.g = g;
        simt::launch(2, 128, 0, [&] { seeded_kernel<2>(gv, queries, seed_pos.data(), nq, out_idx); });
    } else {
        GridView<3, uint32_t, uint32_t> gv;
        gv.x = x.data(); gv.y = y.data(); gv.z = z.data(); gv.idx = idx.data(); gv.cstart = cstart.data(); gv.g = g;
        simt::launch(2, 128, 0, [&] { seeded_kernel<3>(gv, queries, seed_pos.data(), nq, out_idx); });
    }
    return 0;
}

extern "C" void simt_nn_counters(unsigned long long* out6, int reset)
{
    NnCounters& c = nn_counters();
    out6[0] = c.q; out6[1] = c.rows; out6[2] = c.spans; out6[3] = c.cand; out6[4] = c.pass2; out6[5] = c.exact;
    if (reset) c = NnCounters{};
}

extern "C" unsigned long long simt_yield_count() { return simt::st().yields; }
