// emul.cpp — runs the product's kernel headers on the lock-step CTA emulator (cpu_simt.h).
// TEST INFRASTRUCTURE: built into tests/simt/_build/libicp_simt.so by tests/simt/Makefile and
// loaded only by the CPU-only tests.  Never shipped, never a fallback.
#define ICP_SIMT_EMULATE 1
#include "icp_lean.cuh"
#include "icp_stream.cuh"
#include "icp_stats_kernel.cuh"

#include <vector>

using namespace icp;

// ordered: 0 warp-tree (icp_fused.cuh), 1 ordered sums, 2 warp-tree on the icp_lean.cuh schedule
template <int THREADS, int PPT, int DIM>
static void run_fused(FusedParams p, unsigned grid, int ordered)
{
    if (ordered == 2) {
        if constexpr (THREADS >= 64) {
            const LeanLayout LL = lean_layout(DIM, THREADS * PPT, p.mcap, p.cell_budget, false);
            simt::launch(grid, THREADS, LL.total, [&] { icp_lean_kernel<THREADS, PPT, DIM, false, 1>(p); });
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

// Same argument meaning as icp_b200_batch.  variant: 0 -> 128 threads x 3 points, 1 -> 128 x 9,
// 2 -> 64 x 2 (tiny), 3 -> 32 x 4 (single warp).
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
    const FusedLayout L = fused_layout(dim, 128 * 9, mcap, cell_budget, false, true);
    std::vector<unsigned char> ws((size_t)grid * (L.nn_bytes + 64));
    p.workspace = ws.data();
    p.ws_stride = L.nn_bytes + 64 - ((L.nn_bytes + 64) % 16);
    if (dim == 2) {
        if (variant == 0) { p.ncap = 384; run_fused<128, 3, 2>(p, grid, ordered); }
        else if (variant == 1) { p.ncap = 1152; run_fused<128, 9, 2>(p, grid, ordered); }
        else if (variant == 2) { p.ncap = 128; run_fused<64, 2, 2>(p, grid, ordered); }
        else { p.ncap = 128; run_fused<32, 4, 2>(p, grid, ordered); }
    } else {
        if (variant == 0) { p.ncap = 384; run_fused<128, 3, 3>(p, grid, ordered); }
        else if (variant == 1) { p.ncap = 1152; run_fused<128, 9, 3>(p, grid, ordered); }
        else if (variant == 2) { p.ncap = 128; run_fused<64, 2, 3>(p, grid, ordered); }
        else { p.ncap = 128; run_fused<32, 4, 3>(p, grid, ordered); }
    }
    return err;
}

// Streaming (large-scan) kernel on the emulator; threads = 32, 64 or 128.
template <int THREADS, int DIM>
static void run_stream(StreamParams sp, unsigned grid)
{
    simt::launch(grid, THREADS, kStreamShared, [&] { icp_stream_kernel<THREADS, DIM, false>(sp); });
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
    const StreamLayout L = stream_layout(dim, sp.n_cap, sp.m_cap, sp.c_cap, false);
    std::vector<unsigned char> ws((size_t)grid * L.total + 256, 0xCD);
    p.workspace = ws.data();
    p.ws_stride = L.total;
    if (dim == 2) {
        if (threads == 32) run_stream<32, 2>(sp, grid);
        else if (threads == 64) run_stream<64, 2>(sp, grid);
        else run_stream<128, 2>(sp, grid);
    } else {
        if (threads == 32) run_stream<32, 3>(sp, grid);
        else if (threads == 64) run_stream<64, 3>(sp, grid);
        else run_stream<128, 3>(sp, grid);
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

extern "C" unsigned long long simt_yield_count() { return simt::st().yields; }
