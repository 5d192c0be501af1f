// icp_stats_kernel.cuh — stats::calc_stddevs (src/stats.rs:49-60) as a standalone kernel: one CTA
// per residual set.  Exposes the device selection (icp_cta.cuh) through the C ABI so that the
// parity tests can drive it with arbitrary inputs.
#pragma once
#include "icp_cta.cuh"

namespace icp {

struct StatsParams {
    const double* residuals;   // AoS (x y)
    const uint32_t* off;       // [B+1]
    double* out_sigma;         // [B*2]
    double* out_median;        // [B*2] nullable
    uint32_t B;
    int cap;                   // points a CTA can hold
    int mode;                  // 0: radix map only; 1: second evaluation re-using the first one's
                               //    result as the estimate (exercises the value-uniform map)
};

constexpr int kStatsThreads = 128;
constexpr int kStatsPpt = 24;   // keys per thread the selection can hold: sets of up to 3072 points

ICP_HD uint32_t stats_smem_bytes(int cap)
{
    return (uint32_t)(16u * (uint32_t)cap + 16u + sizeof(SelShared));
}

template <int THREADS, int PPT>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(THREADS, 1) calc_stddevs_kernel(StatsParams p)
{
    ICP_SHARED_DYN(smem);
    double* r2 = reinterpret_cast<double*>(smem);
    SelShared* sel = reinterpret_cast<SelShared*>(smem + ((16u * (uint32_t)p.cap + 15u) & ~15u));
    const int tid = (int)threadIdx.x;
    for (uint32_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        const int n = (int)(p.off[b + 1] - p.off[b]);
        const double* res = p.residuals + (size_t)p.off[b] * 2;
        for (int q = tid; q < 2 * kSelHistPad; q += THREADS) (&sel->hist[0][0])[q] = 0u;
        if (n < 1 || n > p.cap) {   // None (stats.rs:15-17) / does not fit
            if (tid < 2) {
                p.out_sigma[(size_t)b * 2 + tid] = -1.0;
                if (p.out_median) p.out_median[(size_t)b * 2 + tid] = 0.0;
            }
            __syncthreads();
            continue;
        }
        for (int q = tid; q < 2 * n; q += THREADS) r2[q] = res[q];
        __syncthreads();
        int par = 0;
        double med[2], mad[2];
        KeysInShared ks;
        ks.v = reinterpret_cast<const double2*>(r2);
        bool ev[2] = {false, false};
        float elo[2] = {0.f, 0.f}, esc[2] = {0.f, 0.f};
        const double zero2[2] = {0., 0.};
        par = block_median2<THREADS, PPT>(sel, ks, n, 0, zero2, ev, elo, esc, par, med);
        par = block_median2<THREADS, PPT>(sel, ks, n, 1, med, ev, elo, esc, par, mad);
        if (p.mode == 1) {
            double med2[2], mad2[2];
            for (int a = 0; a < 2; ++a) {
                const double s = kPpf34 * mad[a];
                ev[a] = sel_fast_params(med[a] - 3.0 * s, 6.0 * s, &elo[a], &esc[a]);
            }
            par = block_median2<THREADS, PPT>(sel, ks, n, 0, zero2, ev, elo, esc, par, med2);
            for (int a = 0; a < 2; ++a) ev[a] = ev[a] && sel_fast_params(0.0, 3.0 * kPpf34 * mad[a], &elo[a], &esc[a]);
            par = block_median2<THREADS, PPT>(sel, ks, n, 1, med2, ev, elo, esc, par, mad2);
            for (int a = 0; a < 2; ++a) { med[a] = med2[a]; mad[a] = mad2[a]; }
        }
        if (tid < 2) {
            p.out_sigma[(size_t)b * 2 + tid] = kPpf34 * mad[tid];
            if (p.out_median) p.out_median[(size_t)b * 2 + tid] = med[tid];
        }
        __syncthreads();
    }
}

}  // namespace icp
