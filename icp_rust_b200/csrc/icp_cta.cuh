// icp_cta.cuh — CTA-cooperative building blocks shared by the ICP kernels:
//   * warp / block reductions
//   * exact order-statistic selection over shared-memory keys (the device counterpart of
//     stats::mutable_median, src/stats.rs:11-28)
//
// All warp collectives are issued by fully converged warps with the full mask.
#pragma once
#include "icp_math.cuh"

namespace icp {

// @region warp_utils
ICP_DEV int lane_id() { return (int)(threadIdx.x & 31u); }
ICP_DEV int warp_id() { return (int)(threadIdx.x >> 5); }

ICP_DEV double shfl_f64(double v, int src)
{
    uint64_t b = f64_to_bits(v);
    uint32_t lo = __shfl_sync(ICP_FULL_MASK, (uint32_t)b, src);
    uint32_t hi = __shfl_sync(ICP_FULL_MASK, (uint32_t)(b >> 32), src);
    return bits_to_f64(((uint64_t)hi << 32) | lo);
}
ICP_DEV double shfl_xor_f64(double v, int m)
{
    uint64_t b = f64_to_bits(v);
    uint32_t lo = __shfl_xor_sync(ICP_FULL_MASK, (uint32_t)b, m);
    uint32_t hi = __shfl_xor_sync(ICP_FULL_MASK, (uint32_t)(b >> 32), m);
    return bits_to_f64(((uint64_t)hi << 32) | lo);
}
ICP_DEV uint64_t shfl_xor_u64(uint64_t b, int m)
{
    uint32_t lo = __shfl_xor_sync(ICP_FULL_MASK, (uint32_t)b, m);
    uint32_t hi = __shfl_xor_sync(ICP_FULL_MASK, (uint32_t)(b >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}
ICP_DEV double warp_min_f64(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor_f64(v, m));
    return v;
}
ICP_DEV double warp_max_f64(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor_f64(v, m));
    return v;
}
ICP_DEV uint64_t warp_sum_u64(uint64_t v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_u64(v, m);
    return v;
}
ICP_DEV uint32_t warp_sum_u32(uint32_t v) { return __reduce_add_sync(ICP_FULL_MASK, v); }

// inclusive prefix sum across the warp
ICP_DEV uint32_t warp_inclusive_scan_u32(uint32_t v)
{
    const int lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(ICP_FULL_MASK, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// Exact selection of order statistics, two key columns ("axes") at a time.
//
// Keys: column a of point i is   f(r2[2*i + a])   with f = identity (median of residuals,
// stats.rs:31) or f = |. - centre_a| (median of absolute deviations, stats.rs:35-36).
//
// Method: one histogram pass over a MONOTONE bin map with explicit UNDER / OVER bins; the bin
// holding the wanted rank is gathered into a <=32-entry list and ranked by one warp.  Because
// the map is monotone, the result is the exact order statistic whatever the bin geometry is;
// geometry only decides speed.  Two maps:
//   FAST  : value-uniform bins around an estimate (previous evaluation's median / sigma);
//   RADIX : 8-bit digits of the order-preserving u64 key under a fixed prefix, MSB first — at
//           most 8 passes, cannot fail (all keys equal once the prefix covers 64 bits).
// FAST is tried once when an estimate exists; any miss falls through to RADIX.
// ---------------------------------------------------------------------------------------------
constexpr int kSelBins = 256;            // regular bins; +2 for UNDER (0) and OVER (257)
constexpr int kSelHist = kSelBins + 2;
constexpr int kSelHistPad = 288;         // 32 lanes x 9 bins, zero padded
constexpr int kSelCap = 32;              // gathered keys ranked by one warp

struct SelBinSpec {
    int mode;        // 0 = FAST (value-uniform), 1 = RADIX
    float lo;        // FAST: value of the left edge of bin 1
    float scale;     // FAST: bins per unit
    uint64_t pfx;    // RADIX: required value of key >> (shift + 8)   (unused when shift == 56)
    int shift;       // RADIX: digit = (key >> shift) & 0xff
};

struct SelDecision {   // written by the locating warp, read by every thread
    int action;        // 0 = gather bins [b_lo, b_hi]; 1 = next radix pass; 2 = miss (restart
                       // in RADIX); 3 = value known (radix prefix exhausted); 4 = FAST bins
                       // [b_lo, b_hi] hold the ranks but too many keys: zoom the map into them
    int b_lo, b_hi;
    uint32_t below;    // number of keys in bins < b_lo
    uint32_t count;    // number of keys in bins [b_lo, b_hi]
};

struct SelShared {
    // [buffer parity][bin]: count of axis 0 in the low half-word, of axis 1 in the high half-word
    // (at most 65535 keys per axis, so the halves never carry into each other)
    uint32_t hist[2][kSelHistPad];
    uint32_t lcnt[2];
    alignas(16) double list[2][kSelCap];   // 16-byte aligned: the service warps read keys in pairs
    SelDecision dec[2];          // written by the axis' service warp, read by every thread
    double out_lo[2], out_hi[2];
};

// FAST bins in binary32: conversion, subtraction of a constant, multiplication by a positive
// constant, clamping and floor are each monotone non-decreasing, so the composition is a
// monotone map of the binary64 key — all that exactness needs.  (A key range too fine for
// binary32 merely crowds one bin; the caller then falls back to RADIX.)
// @region sel_bin
ICP_DEV int sel_bin(const SelBinSpec& s, double v)
{
    if (s.mode == 0) {
        float t = (__double2float_rn(v) - s.lo) * s.scale;
        t = fminf(fmaxf(t, -1.0f), (float)kSelBins);
        return __float2int_rd(t) + 1;  // 0 .. kSelBins+1
    }
    uint64_t key = f64_order_key(v);
    if (s.shift < 56) {
        uint64_t top = key >> (s.shift + 8);
        if (top < s.pfx) return 0;
        if (top > s.pfx) return kSelBins + 1;
    }
    return 1 + (int)((key >> s.shift) & 0xffu);
}

ICP_DEV double sel_key(double r, int absdev, double centre)
{
    return absdev ? fabs(r - centre) : r;
}

// Locate the bin holding rank `k` (0-based) and, if need_hi, rank k+1, in hist[0..kSelHist).
// Executed by one full warp; every lane returns the same decision.
// @region sel_locate
struct SelScan {       // one lane's share of the packed histogram and its warp prefix sums
    uint32_t c[9];     // packed counts of bins lane*9 .. lane*9+8
    uint32_t incl;     // packed inclusive prefix over lanes
    uint32_t tot;
};

ICP_DEV SelScan sel_scan(const uint32_t* hist)
{
    const int lane = lane_id();
    SelScan sc;
    sc.tot = 0;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        sc.c[q] = hist[lane * 9 + q];  // padded with zeros beyond kSelHist
        sc.tot += sc.c[q];
    }
    sc.incl = warp_inclusive_scan_u32(sc.tot);
    return sc;
}

ICP_DEV SelDecision sel_locate(const SelScan& sc, int axis, uint32_t k, bool need_hi, const SelBinSpec& spec)
{
    const int lane = lane_id();
    const int sh = axis * 16;
    uint32_t c[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) c[q] = (sc.c[q] >> sh) & 0xffffu;
    const uint32_t tot = (sc.tot >> sh) & 0xffffu;
    const uint32_t incl = (sc.incl >> sh) & 0xffffu;
    const uint32_t excl = incl - tot;
    SelDecision d;
    // rank k
    uint32_t vote = __ballot_sync(ICP_FULL_MASK, incl > k);
    int owner = __ffs((int)vote) - 1;
    int bin = 0;
    uint32_t below = 0;
    if (lane == owner) {
        uint32_t run = excl;
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            if (run + c[q] > k) { bin = lane * 9 + q; below = run; break; }
            run += c[q];
        }
    }
    bin = __shfl_sync(ICP_FULL_MASK, bin, owner);
    below = __shfl_sync(ICP_FULL_MASK, below, owner);
    d.b_lo = bin;
    d.b_hi = bin;
    d.below = below;
    if (need_hi) {
        uint32_t k2 = k + 1;
        uint32_t vote2 = __ballot_sync(ICP_FULL_MASK, incl > k2);
        int owner2 = __ffs((int)vote2) - 1;
        int bin2 = 0;
        if (lane == owner2) {
            uint32_t run = excl;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                if (run + c[q] > k2) { bin2 = lane * 9 + q; break; }
                run += c[q];
            }
        }
        d.b_hi = __shfl_sync(ICP_FULL_MASK, bin2, owner2);
    }
    // count keys in [b_lo, b_hi]
    uint32_t part = 0;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        int b = lane * 9 + q;
        if (b >= d.b_lo && b <= d.b_hi) part += c[q];
    }
    d.count = warp_sum_u32(part);
    const bool regular = d.b_lo >= 1 && d.b_hi <= kSelBins;
    if (regular && d.count <= (uint32_t)kSelCap) d.action = 0;
    else if (spec.mode == 0) d.action = regular ? 4 : 2;
    else if (d.b_lo != d.b_hi) d.action = 2;          // only reachable with need_hi in RADIX: not used
    else if (spec.shift == 0) d.action = 3;           // every surviving key is identical
    else d.action = 1;
    return d;
}

// Rank the gathered list (cnt <= kSelCap entries) with one warp: returns the entries of rank
// kp and kp+1 (the latter only meaningful when kp+1 < cnt).
// @region sel_finish
ICP_DEV void sel_finish(const double* list, uint32_t cnt, uint32_t kp, double* v0, double* v1)
{
    const int lane = lane_id();
    const bool have = (uint32_t)lane < cnt;
    const double mine = have ? list[lane] : 0.0;
    uint32_t rank = 0;
    for (uint32_t m = 0; m < cnt; ++m) {
        double o = list[m];
        rank += (o < mine || (o == mine && m < (uint32_t)lane)) ? 1u : 0u;
    }
    uint32_t vote0 = __ballot_sync(ICP_FULL_MASK, have && rank == kp);
    uint32_t vote1 = __ballot_sync(ICP_FULL_MASK, have && rank == kp + 1);
    int o0 = __ffs((int)vote0) - 1;
    int o1 = __ffs((int)vote1) - 1;
    *v0 = shfl_f64(mine, o0 < 0 ? 0 : o0);
    *v1 = shfl_f64(mine, o1 < 0 ? 0 : o1);
}

// Per-axis selection state, kept redundantly (and identically) by every thread of the CTA.
// @region sel_state
struct SelAxisState {
    SelBinSpec spec;
    int phase;       // 0 = searching rank k (+k+1 when FAST), 1 = searching rank k+1 (RADIX), 2 = done
    int zooms;       // FAST refinements done so far
    double v_lo, v_hi;
};
constexpr int kSelMaxZooms = 2;

ICP_DEV void sel_start_radix(SelAxisState& st, int which)
{
    st.spec.mode = 1;
    st.spec.pfx = 0;
    st.spec.shift = 56;
    st.phase = which;
}

// FAST-map parameters from an estimate of centre and spread of the keys: kSelBins bins over
// [lo, lo + width).  Returns false when the binary32 map would be degenerate.
ICP_DEV bool sel_fast_params(double lo, double width, float* lo32, float* scale32)
{
    const double sc = (double)kSelBins / width;
    *lo32 = __double2float_rn(lo);
    *scale32 = __double2float_rn(sc);
    return width > 0.0 && sc < 1e30 && sc > 1e-30 && fabs(lo) < 1e30;
}

// Key sources: where block_median2 reads the (x, y) value pair of point i = tid + kk * THREADS.
struct KeysInShared {            // AoS double2 array in shared memory
    const double2* v;
    ICP_DEV_MEMBER double2 get(int /*kk*/, int i) const { return v[i]; }
};
template <int PPT>
struct KeysInRegisters {         // the calling thread's own points, held in registers
    const double2* v;            // v[kk]
    ICP_DEV_MEMBER double2 get(int kk, int /*i*/) const { return v[kk]; }
};

// Medians of the two key columns.  n >= 1; n <= THREADS * PPT, or PPT == 0 for any n <= 65535
// (keys are then visited with a run-time loop and bin numbers are recomputed when gathering).  est_valid[a] enables the FAST map
// for axis a with value range [est_lo[a], est_lo[a] + kSelBins / est_scale[a]).
// On return med[a] = median as stats.rs:11-28 defines it ((b + c) / 2 for even n).
// `hist_parity` names the histogram buffer that is already zeroed on entry; on exit the
// returned parity names a zeroed buffer again.  Must be called by all threads of the CTA.
//
// Round structure: histogram pass | barrier | warp a locates axis a's bins in the shared histogram |
// barrier | all threads gather those bins' keys (cached bin numbers) into a short list | barrier |
// warp a ranks list a | barrier.  The CTA is instruction-issue bound with two or three CTAs per SM,
// so the serial steps are done once (by one warp per axis) rather than redundantly by every warp.
template <int THREADS, int PPT, class KeySrc>
ICP_DEV int block_median2(SelShared* sh, const KeySrc& keys, int n, int absdev, const double centre[2],
                          const bool est_valid[2], const float est_lo[2], const float est_scale[2],
                          int hist_parity, double med[2])
{
    const int tid = (int)threadIdx.x;
    const int lane = lane_id();
    const bool need_hi = (n % 2) == 0;
    const uint32_t k = need_hi ? (uint32_t)(n / 2 - 1) : (uint32_t)(n / 2);
    SelAxisState st[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        st[a].v_lo = 0.;
        st[a].v_hi = 0.;
        st[a].zooms = 0;
        if (est_valid[a]) {
            st[a].spec.mode = 0;
            st[a].spec.lo = est_lo[a];
            st[a].spec.scale = est_scale[a];
            st[a].spec.pfx = 0;
            st[a].spec.shift = 56;
            st[a].phase = 0;
        } else {
            sel_start_radix(st[a], 0);
            st[a].spec.lo = 0.f;
            st[a].spec.scale = 0.f;
        }
    }
    int par = hist_parity;
    while (st[0].phase != 2 || st[1].phase != 2) {
        uint32_t* h = sh->hist[par];
// @region sel_hist_pass
        // ---- histogram pass (also clears the other buffer and the list counters)
        for (int q = tid; q < kSelHistPad; q += THREADS) sh->hist[par ^ 1][q] = 0u;
        if (tid < 2) sh->lcnt[tid] = 0u;
        // this round's bin of each of the thread's keys (axis 0 | axis 1 << 16), cached for the
        // gather pass when the per-thread point count is a compile-time constant
        uint32_t bins[PPT > 0 ? PPT : 1];
        {
            uint32_t under0 = 0, over0 = 0, under1 = 0, over1 = 0;
            auto bin_point = [&](int kk, int i) -> uint32_t {
                const double2 rr = keys.get(kk, i);
                int b0 = 0xffff, b1 = 0xffff;
                if (st[0].phase != 2) {
                    b0 = sel_bin(st[0].spec, sel_key(rr.x, absdev, centre[0]));
                    if (b0 == 0) ++under0;
                    else if (b0 == kSelBins + 1) ++over0;
                    else atomicAdd(&h[b0], 1u);
                }
                if (st[1].phase != 2) {
                    b1 = sel_bin(st[1].spec, sel_key(rr.y, absdev, centre[1]));
                    if (b1 == 0) ++under1;
                    else if (b1 == kSelBins + 1) ++over1;
                    else atomicAdd(&h[b1], 0x10000u);
                }
                return (uint32_t)b0 | ((uint32_t)b1 << 16);
            };
            if (PPT > 0) {
#pragma unroll
                for (int kk = 0; kk < (PPT > 0 ? PPT : 1); ++kk) {
                    const int i = tid + kk * THREADS;
                    bins[kk] = (i < n) ? bin_point(kk, i) : 0xffffffffu;
                }
            } else {
                for (int i = tid, kk = 0; i < n; i += THREADS, ++kk) bin_point(kk, i);
                bins[0] = 0u;
            }
            under0 = warp_sum_u32(under0);
            over0 = warp_sum_u32(over0);
            under1 = warp_sum_u32(under1);
            over1 = warp_sum_u32(over1);
            if (lane == 0) {
                if (under0 | under1) atomicAdd(&h[0], under0 | (under1 << 16));
                if (over0 | over1) atomicAdd(&h[kSelBins + 1], over0 | (over1 << 16));
            }
        }
        __syncthreads();
// @region sel_locate_call
        // ---- locate: warp a serves axis a (a single-warp CTA serves both)
        {
            const int warp = warp_id();
            if (warp < 2) {
                const SelScan sc = sel_scan(h);
                for (int a = (THREADS >= 64 ? warp : 0); a < 2; a += (THREADS >= 64 ? 2 : 1)) {
                    if (st[a].phase == 2) continue;
                    const SelDecision d = sel_locate(sc, a, k + (uint32_t)(st[a].phase == 1 ? 1 : 0),
                                                     (st[a].spec.mode == 0) && need_hi, st[a].spec);
                    if (lane == 0) sh->dec[a] = d;
                }
            }
        }
        __syncthreads();
        SelDecision d0 = sh->dec[0], d1 = sh->dec[1];
        const bool g0 = st[0].phase != 2 && d0.action == 0;
        const bool g1 = st[1].phase != 2 && d1.action == 0;
        double fin_lo[2] = {0., 0.}, fin_hi[2] = {0., 0.};
        if (g0 || g1) {
// @region sel_gather
            // ---- gather the target bins (cached bin numbers; keys are re-read only on a hit)
            if (PPT > 0) {
#pragma unroll
                for (int kk = 0; kk < (PPT > 0 ? PPT : 1); ++kk) {
                    const int i = tid + kk * THREADS;
                    const int b0 = (int)(bins[kk] & 0xffffu), b1 = (int)(bins[kk] >> 16);
                    const bool hit0 = g0 && b0 >= d0.b_lo && b0 <= d0.b_hi;   // 0xffff marks "no key"
                    const bool hit1 = g1 && b1 >= d1.b_lo && b1 <= d1.b_hi;
                    if (hit0 || hit1) {
                        const double2 rr = keys.get(kk, i);
                        if (hit0) sh->list[0][atomicAdd(&sh->lcnt[0], 1u)] = sel_key(rr.x, absdev, centre[0]);
                        if (hit1) sh->list[1][atomicAdd(&sh->lcnt[1], 1u)] = sel_key(rr.y, absdev, centre[1]);
                    }
                }
            } else {
                for (int i = tid, kk = 0; i < n; i += THREADS, ++kk) {
                    const double2 rr = keys.get(kk, i);
                    if (g0) {
                        const double v = sel_key(rr.x, absdev, centre[0]);
                        const int b = sel_bin(st[0].spec, v);
                        if (b >= d0.b_lo && b <= d0.b_hi) sh->list[0][atomicAdd(&sh->lcnt[0], 1u)] = v;
                    }
                    if (g1) {
                        const double v = sel_key(rr.y, absdev, centre[1]);
                        const int b = sel_bin(st[1].spec, v);
                        if (b >= d1.b_lo && b <= d1.b_hi) sh->list[1][atomicAdd(&sh->lcnt[1], 1u)] = v;
                    }
                }
            }
            __syncthreads();
// @region sel_finish_call
            // ---- rank the lists: warp a serves axis a
            {
                const int warp = warp_id();
                if (warp < 2) {
                    for (int a = (THREADS >= 64 ? warp : 0); a < 2; a += (THREADS >= 64 ? 2 : 1)) {
                        if (!(a == 0 ? g0 : g1)) continue;
                        const SelDecision d = a == 0 ? d0 : d1;
                        double v0, v1;
                        sel_finish(sh->list[a], d.count, k + (uint32_t)(st[a].phase == 1 ? 1 : 0) - d.below, &v0, &v1);
                        if (lane == 0) { sh->out_lo[a] = v0; sh->out_hi[a] = v1; }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 2; ++a) { fin_lo[a] = sh->out_lo[a]; fin_hi[a] = sh->out_hi[a]; }
        }
// @region sel_state_update
        // ---- uniform state update
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (st[a].phase == 2) continue;
            const SelDecision d = a == 0 ? d0 : d1;
            const uint32_t kq = k + (uint32_t)(st[a].phase == 1 ? 1 : 0);
            if (d.action == 0) {
                const double v0 = fin_lo[a], v1 = fin_hi[a];
                const bool hi_inside = (kq - d.below) + 1 < d.count;
                if (st[a].phase == 0) {
                    st[a].v_lo = v0;
                    if (!need_hi) st[a].phase = 2;
                    else if (hi_inside) { st[a].v_hi = v1; st[a].phase = 2; }
                    else sel_start_radix(st[a], 1);  // rank k+1 lies in a later bin
                } else {
                    st[a].v_hi = v0;
                    st[a].phase = 2;
                }
            } else if (d.action == 4 && st[a].zooms < kSelMaxZooms) {
                // FAST found the bins but they are crowded (large n): same map, 256x finer, over
                // exactly those bins.  Still monotone; keys outside land in UNDER / OVER and keep
                // the rank bookkeeping right.  If binary32 cannot resolve the range the next
                // round misses and RADIX takes over.
                const float width = (float)(d.b_hi - d.b_lo + 1) / st[a].spec.scale;
                st[a].spec.lo = st[a].spec.lo + (float)(d.b_lo - 1) / st[a].spec.scale;
                st[a].spec.scale = (float)kSelBins / width;
                st[a].zooms += 1;
                if (!(st[a].spec.scale < 1e30f) || !(width > 0.f)) sel_start_radix(st[a], st[a].phase);
            } else if (d.action == 2 || d.action == 4) {
                sel_start_radix(st[a], st[a].phase);
            } else {
                // RADIX: descend into digit (b_lo - 1)
                const uint64_t digit = (uint64_t)(d.b_lo - 1);
                const uint64_t npfx = (st[a].spec.shift == 56) ? digit : ((st[a].spec.pfx << 8) | digit);
                if (d.action == 3) {
                    const double v = f64_from_order_key(npfx);
                    if (st[a].phase == 0) {
                        st[a].v_lo = v;
                        // identical keys fill this bin: rank k+1 is the same value iff it is inside
                        if (!need_hi) st[a].phase = 2;
                        else if ((kq - d.below) + 1 < d.count) { st[a].v_hi = v; st[a].phase = 2; }
                        else sel_start_radix(st[a], 1);
                    } else {
                        st[a].v_hi = v;
                        st[a].phase = 2;
                    }
                } else {
                    st[a].spec.pfx = npfx;
                    st[a].spec.shift -= 8;
                }
            }
        }
        par ^= 1;
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) med[a] = need_hi ? (st[a].v_lo + st[a].v_hi) / 2. : st[a].v_lo;
    return par;
}

// ---------------------------------------------------------------------------------------------
// Sums 16 values per lane across the warp with 16 (instead of 80) 64-bit exchanges: at every
// butterfly step the partners split the value set, each keeping (and summing) one half.  On
// return every lane holds the warp total of value number warp_reduce16_owner(lane) in v[0];
// lanes 2t and 2t+1 hold the same value.  The summation tree is fixed: results do not depend on
// scheduling.
// @region warp_reduce16
ICP_DEV int warp_reduce16_owner(int lane)
{
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}
ICP_DEV void warp_reduce16(double v[16])
{
    const int lane = lane_id();
#pragma unroll
    for (int half = 8, msk = 16; half >= 1; half >>= 1, msk >>= 1) {
        const bool upper = (lane & msk) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const double send = upper ? v[j] : v[j + half];
            const double keep = upper ? v[j + half] : v[j];
            v[j] = keep + shfl_xor_f64(send, msk);
        }
    }
    v[0] = v[0] + shfl_xor_f64(v[0], 1);
}

}  // namespace icp
