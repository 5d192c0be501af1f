// icp_lean.cuh — the whole ICP loop of one scan pair inside one CTA, warp-tree sums, with the
// robust-scale selection served on a short path.
//
// Same algorithm, helpers and results as the warp-tree mode of icp_fused.cuh (Icp2d::estimate
// src/lib.rs:105-130, Icp3d::estimate src/lib.rs:148-173, estimate_transform lib.rs:59-84,
// weighted_gauss_newton_update lib.rs:218-261, calc_stddevs stats.rs:49-60); what differs is how a
// Gauss-Newton evaluation is scheduled on the CTA.  Per evaluation:
//
//   all threads   residuals r = Ti*s - d (lib.rs:230-234) -> shared rows; FAST bin of r per axis
//                 (value-uniform bins around the previous evaluation's median, icp_cta.cuh) ->
//                 packed two-axis histogram + the point's bin numbers in shared memory
//   named barrier the two service warps wait for the histogram; every other warp only ARRIVES and
//                 goes on to its partial sums (below), which need neither sigma nor the medians
//   warp a        axis a: prefix-scan the histogram, find the bin(s) holding the median ranks,
//                 collect those points' keys by scanning the bin numbers, rank them -> exact median
//                 ... and, from the SAME histogram, the exact MAD: |r - median| <= t is the window
//                 [median - t, median + t] of r, so the rank's candidates sit in the two bins the
//                 window's ends reach (lean_serve_mad; verified before it is trusted); sigma =
//                 1.4826*MAD, 1/sigma and the next evaluation's bin maps are published by the warp
//   all threads   (the service warps after serving) Huber weights and the FACTORISED sums of
//                 lean_accumulate_point: five per axis + the cost, reduced by a fixed warp tree
//   barrier
//   (only when lean_serve_mad declines: all threads bin |r - median| into a second histogram,
//    barrier, warp a serves it as in stage 0, barrier)
//   warp 0        adds the warps' partials, applies 1/sigma and R, solves, takes the (guarded) break
//                 decisions and publishes the new inner transform
//   barrier
//
// 2 full barriers + 1 arrive per evaluation (4 on the second-histogram path) and no per-thread selection state.  Whenever the short path cannot
// serve a selection (no estimate yet, rank outside the binned range, more than 32 keys in the
// rank's bins) the CTA runs the general selection block_median2 over the shared rows instead, so
// results never depend on which path ran: both return the exact order statistic.
#pragma once
#ifdef ICP_LEAN_TIMING
#include <cstdio>
#endif
#include "icp_fused.cuh"

namespace icp {

// Short-path histogram: kSelBins regular bins (numbers 1..kSelBins), UNDER = 0, OVER = kSelBins+1;
// word h[b] packs the count of axis 0 (low half) and axis 1 (high half).  h[1] is 16-byte aligned
// so that a service lane reads its 8 bins with two 128-bit loads.
constexpr int kLeanHistWords = 264;          // 3 pad + 258 bins + 3 pad
constexpr int kLeanHistBase = 3;
// Buckets: the points of the kLeanWin bins around the bin the previous evaluation's median maps to
// (number kLeanCentre) are remembered per bin, in arrival order (the slot the histogram atomic
// returns), so the service warp fetches the median's candidates without looking at other points.
constexpr int kLeanWin = 128;   // covers median +- 1 sigma: the MAD's bins (+-0.67 sigma) fall inside
constexpr int kLeanCentre = kSelBins / 2 + 1;
constexpr int kLeanWinLo = kLeanCentre - kLeanWin / 2;
constexpr int kLeanCap = 20;   // 10 KB of buckets: three CTAs of the 1152-point instance share an SM

struct LeanSel {
    SelShared s;                 // histograms + lists of the general selection; list[] also ranks the short path
    alignas(16) uint32_t lh[2][kLeanHistWords];   // [stage] short-path histogram
    uint16_t bucket[2][kLeanWin][kLeanCap];    // [axis][bin - kLeanWinLo][slot] -> point index
    double med[2];               // stage 0 result (median of residuals per axis)
    double sigma[2];             // stage 1 result (1.4826 * MAD)
    double g[2];                 // 1 / sigma
    double est_med[2], est_sigma[2];   // last exact median / sigma (steer the bin maps only)
    float lo0[2], sc0[2], lo1[2], sc1[2];      // bin maps of the next evaluation, stage 0 / stage 1
    int valid[2];                // bin maps usable
    int ok[2];                   // the service warp served its axis in the current stage
    int ok1[2];                  // ... and also found the MAD from the stage-0 histogram
};

struct LeanLayout {
    FusedLayout f;               // target arrays, build temporaries, residual rows (gn_r)
    uint32_t bins;               // u32[ncap]: this stage's bin of (axis 0 | axis 1 << 16) per point
    uint32_t pos;                // u32[ncap]: sorted position of the point's match in the previous outer iteration
    uint32_t dmatch;             // shared targets only, double2[ncap]: the matched target point of this outer iteration (xy)
    uint32_t off_sel, off_ctl, total;
};

ICP_HD LeanLayout lean_layout(int dim, int ncap, int mcap, int cell_budget, bool shared_target)
{
    LeanLayout L;
    L.f = fused_layout(dim, ncap, mcap, cell_budget, shared_target, false);
    L.bins = align16(L.f.gn_r + 16u * (uint32_t)ncap);
    L.pos = align16(L.bins + 4u * (uint32_t)ncap);
    // matched points: a per-pair target stays resident in shared memory, so the match is re-read through
    // its sorted position (pos); a shared target lives in HBM, its matches are copied here
    L.dmatch = align16(L.pos + 4u * (uint32_t)ncap);
    uint32_t u = align128(L.dmatch + (shared_target ? 16u * (uint32_t)ncap : 0u));
    if (u < L.f.u_bytes) u = L.f.u_bytes;
    L.off_sel = u;
    L.off_ctl = align16(L.off_sel + (uint32_t)sizeof(LeanSel));
    L.total = align16(L.off_ctl + (uint32_t)sizeof(FusedCtl));
    return L;
}

#ifdef ICP_LEAN_STATS
void lean_stats_record(int stage, int ok, int direct, int b_lo, int b_hi, int count);
#endif
// Value-uniform bin in binary32 (icp_cta.cuh sel_bin, mode 0): each step is monotone
// non-decreasing in v, which is all exactness needs.  0 = UNDER, kSelBins + 1 = OVER.
// @region lean_bin
ICP_DEV uint32_t lean_bin(double v, float lo, float scale)
{
    float t = (__double2float_rn(v) - lo) * scale;
    t = fminf(fmaxf(t, -1.0f), (float)kSelBins);
    return (uint32_t)(__float2int_rd(t) + 1);
}

// histogram + bucket insertion of one key (all threads)
ICP_DEV void lean_insert(uint32_t* h, uint16_t (*bucket)[kLeanCap], uint32_t b, uint32_t unit, int sh, int i)
{
    const uint32_t slot = (atomicAdd(&h[b], unit) >> sh) & 0xffffu;
    const uint32_t wb = b - (uint32_t)kLeanWinLo;
    if (wb < (uint32_t)kLeanWin && slot < (uint32_t)kLeanCap) bucket[wb][slot] = (uint16_t)i;
}

// Bin holding rank `t` (0-based) among the regular bins whose counts the lanes hold (8 per lane,
// inclusive prefix `incl`, lane total `tot`): returns bin number, keys below it, its count.
// All lanes return the same values.  Requires t < total.
// @region lean_locate
ICP_DEV void lean_find_rank(const uint32_t c[8], uint32_t incl, uint32_t tot, uint32_t t, uint32_t* bin,
                            uint32_t* below, uint32_t* count)
{
    const int lane = lane_id();
    const uint32_t vote = __ballot_sync(ICP_FULL_MASK, incl > t);
    const int owner = __ffs((int)vote) - 1;
    uint32_t run = incl - tot, bq = 0, bb = 0, bc = 0;
    bool found = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const bool hit = !found && (run + c[q] > t);
        if (hit) { bq = (uint32_t)q; bb = run; bc = c[q]; found = true; }
        run += c[q];
    }
    *bin = __shfl_sync(ICP_FULL_MASK, (uint32_t)(8 * lane + 1) + bq, owner);
    *below = __shfl_sync(ICP_FULL_MASK, bb, owner);
    *count = __shfl_sync(ICP_FULL_MASK, bc, owner);
}

// Rank of `mine` among list[0..count): keys below it and keys not above it.  Keys are read two at a
// time; the writer pads an odd count with +inf (ranks of finite keys are unaffected).
ICP_DEV void lean_rank(const double* list, uint32_t count, double mine, uint32_t* nlt, uint32_t* nle)
{
    const double2* l2 = reinterpret_cast<const double2*>(list);
    uint32_t a = 0, b = 0;
    for (uint32_t m = 0; m < count; m += 2) {
        const double2 o = l2[m >> 1];
        a += ((o.x < mine) ? 1u : 0u) + ((o.y < mine) ? 1u : 0u);
        b += ((o.x <= mine) ? 1u : 0u) + ((o.y <= mine) ? 1u : 0u);
    }
    *nlt = a; *nle = b;
}
ICP_DEV void lean_pad_list(double* list, uint32_t count, int lane)
{
    if ((uint32_t)lane == count && (count & 1u) && count < (uint32_t)kSelCap) list[count] = 1.0 / 0.0;
}

// One warp serves one axis of one selection stage: returns true and the exact median of the
// axis' keys (stats.rs:11-28: (b + c) / 2 for even n), or false when the short path cannot
// decide (caller falls back to block_median2).  Keys: stage 0 r[i].axis, stage 1
// |r[i].axis - centre|.
// @region lean_serve
ICP_DEV bool lean_serve(const uint32_t* h, const uint16_t (*bucket)[kLeanCap], const uint32_t* bins,
                        const double2* rows, double* list, int n, int axis, int absdev, double centre, double* out,
                        uint32_t* pfx = nullptr)
{
    const int lane = lane_id();
    const int sh = axis * 16;
    const bool need_hi = (n % 2) == 0;
    const uint32_t k = need_hi ? (uint32_t)(n / 2 - 1) : (uint32_t)(n / 2);
    // ---- locate
    uint32_t c[8];
    {
        const uint4 w0 = *reinterpret_cast<const uint4*>(h + 1 + 8 * lane);
        const uint4 w1 = *reinterpret_cast<const uint4*>(h + 5 + 8 * lane);
        c[0] = w0.x; c[1] = w0.y; c[2] = w0.z; c[3] = w0.w; c[4] = w1.x; c[5] = w1.y; c[6] = w1.z; c[7] = w1.w;
    }
    uint32_t tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) tot += c[q];
    uint32_t incl = warp_inclusive_scan_u32(tot);   // packed halves never carry: <= 65535 keys per axis
    const uint32_t under = (h[0] >> sh) & 0xffffu;
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q] = (c[q] >> sh) & 0xffffu;
    tot = (tot >> sh) & 0xffffu;
    incl = ((incl >> sh) & 0xffffu) + under;
    if (pfx) {   // inclusive prefix of the bins for lean_serve_mad: pfx[b] = keys in bins 0..b
        uint32_t run = incl - tot;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            run += c[q];
            pfx[8 * lane + 1 + q] = run;
        }
        if (lane == 0) pfx[0] = under;
    }
    const uint32_t regular_end = __shfl_sync(ICP_FULL_MASK, incl, 31);   // keys in UNDER + regular bins
    const uint32_t k_top = need_hi ? k + 1u : k;
    bool ok = k >= under && k_top < regular_end;
    uint32_t b_lo = 0, below = 0, c_lo = 0, b_hi = 0, c_hi = 0;
    if (ok) {
        lean_find_rank(c, incl, tot, k, &b_lo, &below, &c_lo);
        b_hi = b_lo;
        if (need_hi && k + 1u >= below + c_lo) {   // rank k+1 opens the next non-empty bin
            uint32_t below_hi;
            lean_find_rank(c, incl, tot, k + 1u, &b_hi, &below_hi, &c_hi);
        }
    }
    const uint32_t count = c_lo + c_hi;
    ok = ok && count <= (uint32_t)kSelCap;
    // ---- fetch the candidates: from the buckets, else by scanning the points' bin numbers
    const uint32_t w_lo = b_lo - (uint32_t)kLeanWinLo, w_hi = b_hi - (uint32_t)kLeanWinLo;
    const bool direct = w_lo < (uint32_t)kLeanWin && w_hi < (uint32_t)kLeanWin && c_lo <= (uint32_t)kLeanCap &&
                        c_hi <= (uint32_t)kLeanCap;
#ifdef ICP_LEAN_STATS
    if (lane == 0) lean_stats_record(absdev, ok ? 1 : 0, direct ? 1 : 0, (int)b_lo, (int)b_hi, (int)count);
#endif
    if (!ok) return false;
    double mine = 0.;
    if (direct) {
        if ((uint32_t)lane < count) {
            const int i = (uint32_t)lane < c_lo ? (int)bucket[w_lo][lane] : (int)bucket[w_hi][(uint32_t)lane - c_lo];
            const double2 rr = rows[i];
            mine = sel_key(axis ? rr.y : rr.x, absdev, centre);
            list[lane] = mine;
        }
        lean_pad_list(list, count, lane);
        __syncwarp();
    } else {
        uint32_t cnt = 0;
        const uint32_t lt = (1u << lane) - 1u;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const uint32_t w = (i < n) ? bins[i] : 0xffffffffu;
            const uint32_t b = (w >> sh) & 0xffffu;
            const bool hit = b == b_lo || b == b_hi;    // 0xffff ("no key") never hits
            const uint32_t m = __ballot_sync(ICP_FULL_MASK, hit);
            if (m) {
                if (hit) {
                    const double2 rr = rows[i];
                    list[cnt + (uint32_t)__popc(m & lt)] = sel_key(axis ? rr.y : rr.x, absdev, centre);
                }
                cnt += (uint32_t)__popc(m);
            }
        }
        lean_pad_list(list, count, lane);
        __syncwarp();
        if ((uint32_t)lane < count) mine = list[lane];
    }
    // ---- rank: lt = keys below mine, le = keys not above; the key of rank r satisfies lt <= r < le
    uint32_t nlt = 0, nle = 0;
    lean_rank(list, count, mine, &nlt, &nle);
    const uint32_t kp = k - below;
    const bool have = (uint32_t)lane < count;
    const uint32_t vote0 = __ballot_sync(ICP_FULL_MASK, have && nlt <= kp && kp < nle);
    const uint32_t vote1 = __ballot_sync(ICP_FULL_MASK, have && nlt <= kp + 1u && kp + 1u < nle);
    const double v0 = shfl_f64(mine, __ffs((int)vote0) - 1);
    if (need_hi) {
        const double v1 = shfl_f64(mine, __ffs((int)vote1) - 1);
        *out = (v0 + v1) / 2.;
    } else {
        *out = v0;
    }
    return true;
}

// MAD of one axis from the STAGE-0 histogram (bins of r under the map (lo, scale)), by the warp
// that has just found the exact median `med`: the k-th smallest |r - med| is the smallest t whose
// window [med - t, med + t] holds k + 1 points.  Bins are uniform in value, so as t grows the
// window's ends cross bin edges alternately; state s = 0, 1, 2, ... names the pair of bins (bl, bh)
// the ends lie in.  For the first state whose bins [bl, bh] hold more than the wanted rank(s), the
// points strictly between bl and bh are closer than the answer, the points outside farther, and the
// points of bl and bh are ranked exactly.  The state arithmetic is only a guess: the result is
// returned only if binning the window's ends for the ranked answer reproduces (bl, bh) with slack
// on the safe side — that check alone carries the proof:
//   bound: bin() is monotone.  bin(r) > bl = bin(med - v0 + eps) gives r > med - v0 + eps, i.e.
//   computed |r - med| < v0 (eps = 1e-15 (|med| + t) exceeds every rounding involved); bin(r) < bl =
//   bin(med - v1 - eps) gives r < med - v1 - eps, i.e. computed |r - med| > v1; same on the right.
// Otherwise (ends near a bin edge, crowded or far bins) false: the caller bins |r - med| itself.
// `pfx`: the inclusive bin prefix lean_serve left for this axis (260 words owned by this warp).
// @region lean_serve_mad
ICP_DEV bool lean_serve_mad(const uint16_t (*bucket)[kLeanCap], const double2* rows, double* list,
                            const uint32_t* pfx, int n, int axis, double med, float lo, float scale, double* out)
{
    const int lane = lane_id();
    const bool need_hi = (n % 2) == 0;
    const uint32_t k = need_hi ? (uint32_t)(n / 2 - 1) : (uint32_t)(n / 2);
    const uint32_t k_top = need_hi ? k + 1u : k;
    // pfx[b] = keys in bins 0..b, written by this warp in lean_serve (the __syncwarp()s there order it)
    // ---- the median's bin and its position inside it
    float u = (__double2float_rn(med) - lo) * scale;
    if (!(u >= 0.0f && u < (float)kSelBins)) return false;
    const float uf = floorf(u);
    const int bm = (int)uf + 1;
    const bool left_first = (u - uf) < 0.5f;
    // ---- first state whose bins hold rank k_top
    int s_star = -1;
    uint32_t votes[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int s = 32 * q + lane;
        const int a = (s + 1) >> 1, b = s >> 1;
        const int bl = bm - (left_first ? a : b), bh = bm + (left_first ? b : a);
        const bool in_range = bl >= 1 && bh <= kSelBins;
        const uint32_t total = in_range ? pfx[bh] - pfx[bl - 1] : 0u;
        votes[q] = __ballot_sync(ICP_FULL_MASK, in_range && total > k_top);
    }
#pragma unroll
    for (int q = 3; q >= 0; --q)
        if (votes[q]) s_star = 32 * q + __ffs((int)votes[q]) - 1;
    if (s_star < 0) return false;
    // bins of state s* and of the state after it: the answer is reached in one of the two (by the end
    // of s*+1 the window covers every point of s*'s bins, more than the rank needs)
    const int sa = (s_star + 1) >> 1, sb = s_star >> 1, na = (s_star + 2) >> 1, nb = (s_star + 1) >> 1;
    const int bl = bm - (left_first ? sa : sb), bh = bm + (left_first ? sb : sa);
    const int nl = bm - (left_first ? na : nb), nh = bm + (left_first ? nb : na);
    if (nl < 1 || nh > kSelBins) return false;
    const uint32_t inside = bh > bl ? pfx[bh - 1] - pfx[bl] : 0u;
    // candidate bins, ascending, duplicates with count 0: nl (<= bl), bl, bh (>= bl), nh (>= bh)
    const uint32_t c0 = nl < bl ? pfx[nl] - pfx[nl - 1] : 0u;
    const uint32_t c1 = pfx[bl] - pfx[bl - 1];
    const uint32_t c2 = bh > bl ? pfx[bh] - pfx[bh - 1] : 0u;
    const uint32_t c3 = nh > bh ? pfx[nh] - pfx[nh - 1] : 0u;
    const uint32_t count = c0 + c1 + c2 + c3;
    const uint32_t cap = (uint32_t)kLeanCap;
    if (!(inside <= k && count <= (uint32_t)kSelCap && c0 <= cap && c1 <= cap && c2 <= cap && c3 <= cap &&
          (uint32_t)(nl - kLeanWinLo) < (uint32_t)kLeanWin && (uint32_t)(nh - kLeanWinLo) < (uint32_t)kLeanWin))
        return false;
    // ---- rank the points of those bins on their exact |r - med|
    __syncwarp();   // list[] still holds the median's candidates: every lane has finished reading them
    double mine = 0.;
    if ((uint32_t)lane < count) {
        uint32_t o = (uint32_t)lane;
        int bin = nl;
        if (o >= c0) { o -= c0; bin = bl; if (o >= c1) { o -= c1; bin = bh; if (o >= c2) { o -= c2; bin = nh; } } }
        const int i = (int)bucket[bin - kLeanWinLo][o];
        const double2 rr = rows[i];
        mine = sel_key(axis ? rr.y : rr.x, 1, med);
        list[lane] = mine;
    }
    lean_pad_list(list, count, lane);
    __syncwarp();
    uint32_t nlt = 0, nle = 0;
    lean_rank(list, count, mine, &nlt, &nle);
    const uint32_t kp = k - inside;
    const bool have = (uint32_t)lane < count;
    const uint32_t vote0 = __ballot_sync(ICP_FULL_MASK, have && nlt <= kp && kp < nle);
    const uint32_t vote1 = __ballot_sync(ICP_FULL_MASK, have && nlt <= kp + 1u && kp + 1u < nle);
    if (vote0 == 0u || (need_hi && vote1 == 0u)) return false;
    const double v0 = shfl_f64(mine, __ffs((int)vote0) - 1);
    const double v1 = need_hi ? shfl_f64(mine, __ffs((int)vote1) - 1) : v0;
    // ---- verify against the answer (see bound above): the points strictly between bl and bh are
    // closer than v0, the points beyond nl / nh farther than v1
    const double e0 = (fabs(med) + v0) * 1e-15, e1 = (fabs(med) + v1) * 1e-15;
    const bool good = lean_bin((med - v0) + e0, lo, scale) <= (uint32_t)bl && lean_bin((med + v0) - e0, lo, scale) >= (uint32_t)bh &&
                      lean_bin((med - v1) - e1, lo, scale) >= (uint32_t)nl && lean_bin((med + v1) + e1, lo, scale) <= (uint32_t)nh;
    if (!good) return false;
    *out = need_hi ? (v0 + v1) / 2. : v0;
    return true;
}

template <int THREADS>
ICP_DEV void lean_zero_hist(uint32_t* h)
{
    for (int q = (int)threadIdx.x; q < kSelHistPad; q += THREADS) h[q] = 0u;
}
template <int THREADS>
ICP_DEV void lean_zero_lh(uint32_t* h)
{
    for (int q = (int)threadIdx.x; q < kLeanHistWords; q += THREADS) h[q] = 0u;
}

// publishes the result of a finished calc_stddevs for axis a: sigma, 1/sigma, and the bin maps
// the next evaluation will use: kSelBins bins over median +- 2 sigma, then over MAD +- sigma, so
// that both stages expect their answer in bin kLeanCentre
// @region lean_publish
ICP_DEV void lean_publish(LeanSel* sel, int a, double med, double sigma)
{
    const double g = 1. / sigma;
    sel->sigma[a] = sigma;
    sel->g[a] = g;
    sel->est_med[a] = med;
    sel->est_sigma[a] = sigma;
    // bins/width from the quotient already formed (the maps only steer the selection; any
    // monotone map gives the same exact order statistic): kSelBins/(4 sigma), kSelBins/(2 sigma)
    const double lo0 = med - 2.0 * sigma, sc0 = (0.25 * (double)kSelBins) * g;
    const double lo1 = sigma * (1.0 / kPpf34) - sigma, sc1 = (0.5 * (double)kSelBins) * g;
    sel->lo0[a] = __double2float_rn(lo0);
    sel->sc0[a] = __double2float_rn(sc0);
    sel->lo1[a] = __double2float_rn(lo1);
    sel->sc1[a] = __double2float_rn(sc1);
    sel->valid[a] = (sigma > 0.0 && sc0 < 1e30 && sc1 > 1e-30 && fabs(lo0) < 1e30 && fabs(lo1) < 1e30) ? 1 : 0;
}

// General exact selection over the shared rows (any n, any distribution): block_median2 with its
// FAST / zoom / RADIX maps.  Out of line: it runs once per pair (no estimate yet) and whenever the
// short path cannot decide, and must not cost the short path registers.
// @region lean_general_selection
template <int THREADS, int PPT>
ICP_DEV_NOINLINE void lean_general_selection(LeanSel* sel, const double2* r2v, int n)
{
    const int tid = (int)threadIdx.x;
    uint32_t* h0 = sel->s.hist[0];
    uint32_t* h1 = sel->s.hist[1];
    lean_zero_hist<THREADS>(h0);
    lean_zero_hist<THREADS>(h1);
    __syncthreads();
    double em[2], es[2], sg[2];
    em[0] = sel->est_med[0]; em[1] = sel->est_med[1];
    es[0] = sel->est_sigma[0]; es[1] = sel->est_sigma[1];
    KeysInShared ks;
    ks.v = r2v;
    cta_calc_stddevs<THREADS, PPT>(&sel->s, ks, n, em, es, 0, sg);
    lean_zero_lh<THREADS>(sel->lh[0]);
    lean_zero_lh<THREADS>(sel->lh[1]);
    if (tid < 2) lean_publish(sel, tid, tid ? em[1] : em[0], tid ? sg[1] : sg[0]);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Factorised normal-equation sums.  Row j of a point has J = [Rj0, Rj1, b_j] (lib.rs:176-184) with
// Rj0, Rj1 the same for every point, and its weight is c = w_j * g_j with w_j = drho(r_j^2) of the
// UNSCALED residual (lib.rs:250) and g_j = 1 / sigma_j the same for every point (lib.rs:248).  So
//   J^T W J = sum_j g_j [ Rj0^2 A_j   Rj0 Rj1 A_j   Rj0 B_j ]      J^T W r = sum_j g_j [ Rj0 D_j ]
//                       [     .        Rj1^2 A_j    Rj1 B_j ]                          [ Rj1 D_j ]
//                       [     .            .          C_j   ]                          [   E_j   ]
// with A = sum w, B = sum w b, C = sum w b^2, D = sum w r, E = sum w b r per axis: five sums per axis
// that need neither sigma nor g.  The threads form them while the service warps are still selecting
// the medians, and warp 0 applies g and R once.  Same quantities as lib.rs:242-255 up to summation
// rounding (tree mode: decisions are guarded, transforms agree to ~1e-13).
// acc: 0-4 axis 0 (A B C D E), 5-9 axis 1, 10 Huber cost (lib.rs:47-48), 11-15 unused (zero).
// The two Huber weights and the Huber cost of one point.  (Out of line — to shrink the evaluation loop, whose
// 14k instructions exceed the instruction cache: ncu shows 2.2 warps per issue slot waiting for instructions
// with three CTAs per SM — was measured: same throughput, the quotients and the root are not what is large.)
ICP_DEV void lean_huber3(double rx, double ry, double* w0, double* w1, double* rho)
{
    *w0 = huber_drho_of_residual(rx);
    *w1 = huber_drho_of_residual(ry);
    *rho = huber_rho(rx * rx + ry * ry);
}

// @region lean_accumulate
ICP_DEV void lean_accumulate_point(double acc[16], const Transform& Ti, double sx, double sy, double2 rr, bool valid)
{
    const double a0 = -sy, a1 = sx;   // jacobian(), lib.rs:176-184
    const double b0 = Ti.r00 * a0 + Ti.r01 * a1, b1 = Ti.r10 * a0 + Ti.r11 * a1;
    // (a 20-bit hardware seed + Newton steps in FMA arithmetic instead of the correctly rounded quotient and
    // root was measured: 156k against 166k pairs/s — its dependent FMA chain is longer than the division's)
    double w0r, w1r, rho;
    lean_huber3(rr.x, rr.y, &w0r, &w1r, &rho);
    const double w0 = valid ? w0r : 0., w1 = valid ? w1r : 0.;
    const double wb0 = w0 * b0, wb1 = w1 * b1;
    acc[0] = acc[0] + w0; acc[1] = acc[1] + wb0; acc[2] = acc[2] + wb0 * b0;
    acc[3] = acc[3] + w0 * rr.x; acc[4] = acc[4] + wb0 * rr.x;
    acc[5] = acc[5] + w1; acc[6] = acc[6] + wb1; acc[7] = acc[7] + wb1 * b1;
    acc[8] = acc[8] + w1 * rr.y; acc[9] = acc[9] + wb1 * rr.y;
    acc[10] = acc[10] + (valid ? rho : 0.);
}

// Warp 0: adds the warps' partial sums in warp order, applies g and R (see above; an axis with
// sigma == 0 is skipped as lib.rs:245-247 skips its rows), then solves and decides (tree_decide).
// Ends with a barrier: every thread then reads ctl->bcast.
// @region lean_combine
template <int NW>
ICP_DEV void lean_combine_solve(FusedCtl* ctl, const LeanSel* sel, const Transform& Ti, double prev_error,
                                const GuardBands& gb)
{
    const int lane = lane_id(), warp = warp_id();
    if (warp == 0) {
        double tot = 0.;
        if (lane < 16) {
            tot = ctl->red[0][lane];
            for (int w = 1; w < NW; ++w) tot = tot + ctl->red[w][lane];
        }
        double v[11];
#pragma unroll
        for (int q = 0; q < 11; ++q) v[q] = shfl_f64(tot, q);
        double m00 = 0., m01 = 0., m02 = 0., m11 = 0., m12 = 0., m22 = 0., t0 = 0., t1 = 0., t2 = 0.;
        if (sel->sigma[0] != 0.) {
            const double g = sel->g[0], p = Ti.r00, q = Ti.r01;
            const double gA = g * v[0], gB = g * v[1], gD = g * v[3];
            m00 = (gA * p) * p; m01 = (gA * p) * q; m02 = gB * p; m11 = (gA * q) * q; m12 = gB * q; m22 = g * v[2];
            t0 = gD * p; t1 = gD * q; t2 = g * v[4];
        }
        if (sel->sigma[1] != 0.) {
            const double g = sel->g[1], p = Ti.r10, q = Ti.r11;
            const double gA = g * v[5], gB = g * v[6], gD = g * v[8];
            m00 = m00 + (gA * p) * p; m01 = m01 + (gA * p) * q; m02 = m02 + gB * p;
            m11 = m11 + (gA * q) * q; m12 = m12 + gB * q; m22 = m22 + g * v[7];
            t0 = t0 + gD * p; t1 = t1 + gD * q; t2 = t2 + g * v[9];
        }
        const double jtj[9] = {m00, m01, m02, m01, m11, m12, m02, m12, m22};
        const double jtr[3] = {t0, t1, t2};
        tree_decide(ctl, jtj, jtr, v[10], Ti, prev_error, gb);
    }
    __syncthreads();
}

// Phase timing of CTA 0 (debug builds only, -DICP_LEAN_TIMING): thread 0 accumulates clock64()
// differences between the barriers of an evaluation and prints the per-pair averages.
#ifdef ICP_LEAN_TIMING
#define ICP_LT_DECL long long lt_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long lt_t = 0, lt_n = 0, lt_e = 0;
#define ICP_LT_START lt_t = clock64();
#define ICP_LT_MARK(q) { const long long lt_now = clock64(); lt_acc[q] += lt_now - lt_t; lt_t = lt_now; }
#else
#define ICP_LT_DECL
#define ICP_LT_START
#define ICP_LT_MARK(q)
#endif

// ---------------------------------------------------------------------------------------------
// @region lean_pair_setup
// SVC: warps 0 and 1 are DEDICATED service warps — they own no points, so the medians are being
// served from the moment the histogram is complete while the THREADS - 64 worker threads form the
// factorised sums; capacity (THREADS - 64) * PPT points.  !SVC: every warp owns points and warps
// 0 and 1 add their own sums after serving (capacity THREADS * PPT).
template <int THREADS, int PPT, int DIM, bool SHARED_TARGET, int MINB, bool SVC>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(THREADS, MINB) icp_lean_kernel(FusedParams p)
{
    constexpr int NW = THREADS / 32;
    static_assert(NW <= kMaxWarps, "FusedCtl holds per-warp slots for kMaxWarps warps");
    constexpr int WT0 = SVC ? 64 : 0;            // first worker thread
    constexpr int WTHREADS = THREADS - WT0;      // threads that own points
    constexpr int NCAP = WTHREADS * PPT;
    static_assert(THREADS >= 64 && WTHREADS >= 32, "two service warps (+ at least one worker warp)");
    ICP_SHARED_DYN(smem);
    const LeanLayout LL = lean_layout(DIM, NCAP, p.mcap, p.cell_budget, SHARED_TARGET);
    const FusedLayout& L = LL.f;
    unsigned char* U = smem;
    LeanSel* sel = reinterpret_cast<LeanSel*>(smem + LL.off_sel);
    FusedCtl* ctl = reinterpret_cast<FusedCtl*>(smem + LL.off_ctl);
    double2* r2v = reinterpret_cast<double2*>(U + L.gn_r);
    uint32_t* binv = reinterpret_cast<uint32_t*>(U + LL.bins);
    uint32_t* prev_pos = reinterpret_cast<uint32_t*>(U + LL.pos);
    double2* d2v = reinterpret_cast<double2*>(U + LL.dmatch);
    uint32_t* h0 = sel->lh[0] + kLeanHistBase;
    uint32_t* h1 = sel->lh[1] + kLeanHistBase;

    const int tid = (int)threadIdx.x, lane = lane_id(), warp = warp_id();
    const bool worker = !SVC || warp >= 2;
    const int wtid = tid - WT0;                  // point i = wtid + k * WTHREADS belongs to worker wtid
    // the target point matched to source point i in the current outer iteration (xy, get_xy lib.rs:86-89)
    auto matched = [&](int i) -> double2 {
        if (SHARED_TARGET) return d2v[i];
        const uint32_t pos = prev_pos[i];
        double2 d;
        d.x = reinterpret_cast<const double*>(U + L.nn_x)[pos];
        d.y = reinterpret_cast<const double*>(U + L.nn_y)[pos];
        return d;
    };

    lean_zero_lh<THREADS>(sel->lh[0]);
    lean_zero_lh<THREADS>(sel->lh[1]);
    fused_launch_entry(p);
    __syncthreads();
    ICP_LT_DECL

    for (;;) {
        ICP_LT_START
        if (tid == 0) { ctl->pair = fused_next_pair(p); ctl->knife = 0u; }
        __syncthreads();
        const uint32_t pair = ctl->pair;
        if (pair >= p.pair_end) break;
        const uint32_t spair = p.shared_src ? 0u : pair;   // ICP_B200_FLAG_SHARED_SOURCE: one scan for every pair
        const int n = (int)(p.src_off[spair + 1] - p.src_off[spair]);
        const double* src = p.src + (size_t)p.src_off[spair] * DIM;
        int m;
        const double* dst = nullptr;
        if (SHARED_TARGET) {
            m = (int)p.target.m;
        } else {
            m = (int)(p.dst_off[pair + 1] - p.dst_off[pair]);
            dst = p.dst + (size_t)p.dst_off[pair] * DIM;
        }
        // the outer transform and the checksum live in shared memory (ctl->T6, ctl->H): they are
        // touched once per outer iteration and would otherwise cost every thread 14 registers
        if (tid < 6) ctl->T6[tid] = p.init_T[(size_t)pair * 6 + tid];
        if (tid == 0) ctl->H = 0ull;

        int status = kPairOk;
        if (m == 0 && n > 0 && p.max_iter > 0) status = kPairEmptyTarget;
        if (n > NCAP || (!SHARED_TARGET && m > p.mcap)) status = kPairTooLarge;
        {
            // the reference panics on NaN residuals (stats.rs:12); reject non-finite input up front
            int bad = 0;
            for (int q = tid; q < n * DIM; q += THREADS) bad |= !(fabs(src[q]) <= kF64Max);
            if (!SHARED_TARGET)
                for (int q = tid; q < m * DIM; q += THREADS) bad |= !(fabs(dst[q]) <= kF64Max);
            if (tid < 6) bad |= !(fabs(p.init_T[(size_t)pair * 6 + tid]) <= kF64Max);
            if (__syncthreads_or(bad) && status == kPairOk) status = kPairNonFinite;
        }
        if (status != kPairOk) {
            if (tid == 0) {
#pragma unroll
                for (int q = 0; q < 6; ++q) p.out_T[(size_t)pair * 6 + q] = p.init_T[(size_t)pair * 6 + q];
                if (p.out_status) p.out_status[pair] = status;
                if (p.out_error) atomicMax(p.out_error, status);
                if (p.out_hash) p.out_hash[pair] = 0ull;
                if (p.out_cost) p.out_cost[pair] = 0.;
                if (p.out_inner)
                    for (uint32_t it = 0; it < p.max_iter; ++it)
                        p.out_inner[(size_t)pair * p.max_iter + it] = 0u;
            }
            __syncthreads();
            continue;
        }

        // ---- Icp2d::new / Icp3d::new (lib.rs:97-102,139-144): sorted cell list of the target
        GridGeom geom;
        if (!SHARED_TARGET) {
            if (m > 0) geom = cta_build_grid<THREADS, DIM>(dst, m, U, L, ctl, p.cell_budget);
        } else {
            geom = p.target.g;
        }
        if (tid < 2) {   // no estimate of this pair's residual distribution yet
            sel->valid[tid] = 0;
            sel->est_med[tid] = 0.;
            sel->est_sigma[tid] = 0.;
        }

        const uint64_t corr_base = (p.shared_src ? (uint64_t)pair * (uint64_t)n : (uint64_t)p.src_off[pair]) *
                                   (p.corr_mode == 2 ? p.max_iter : 1u);

        ICP_LT_MARK(0)
        for (uint32_t it = 0; it < p.max_iter; ++it) {
// @region lean_search
            // ---- transform (lib.rs:113-116) + correspondence search (lib.rs:118-124): one inlined
            // copy of the search; queries and matched points go through the thread's own row slots
            double sx[PPT], sy[PPT];
            uint64_t hpart = 0;
            {
                Transform T;   // visible: written before the barrier that ended the previous iteration / the input check
                T.r00 = ctl->T6[0]; T.r01 = ctl->T6[1]; T.r10 = ctl->T6[2]; T.r11 = ctl->T6[3]; T.tx = ctl->T6[4]; T.ty = ctl->T6[5];
                for (int i = tid; i < NCAP; i += THREADS) {
                    double2 q;
                    q.x = 0.; q.y = 0.;
                    if (i < n) transform_apply(T, src[(size_t)i * DIM + 0], src[(size_t)i * DIM + 1], &q.x, &q.y);
                    else if (SHARED_TARGET) d2v[i] = q;   // slots beyond the scan: weight 0 whatever they hold
                    else prev_pos[i] = 0u;
                    r2v[i] = q;
                }
            }
            // (a group of four lanes per query on a shared target — target_nn_search2_group — was measured in
            // this loop: config 3 48.7k against 54.9k scans/s; the stand-alone kernel 4.1 against 6.4 G queries/s)
#pragma unroll 1
            for (int i = tid; i < n; i += THREADS) {
                const double2 q = r2v[i];
                const double pz = (DIM == 3) ? src[(size_t)i * DIM + 2] : 0.0;  // transform_xy keeps z (lib.rs:52-57)
                double2 d;
                int idx;
                // from the second outer iteration on the previous match seeds the search
                const int seed = (it > 0) ? (int)prev_pos[i] : -1;
                if (SHARED_TARGET) {
                    const NnResult<DIM, uint32_t, uint32_t> r = target_nn_search<DIM>(p.target, q.x, q.y, pz, seed);
                    d.x = p.target.x[r.pos]; d.y = p.target.y[r.pos];
                    idx = r.idx;
                    prev_pos[i] = (uint32_t)r.pos;
                } else {
                    GridView<DIM, uint16_t, uint16_t> gv;
                    gv.x = reinterpret_cast<const double*>(U + L.nn_x);
                    gv.y = reinterpret_cast<const double*>(U + L.nn_y);
                    gv.z = reinterpret_cast<const double*>(U + L.nn_z);
                    gv.idx = reinterpret_cast<const uint16_t*>(U + L.nn_idx);
                    gv.cstart = reinterpret_cast<const uint16_t*>(U + L.nn_cell);
                    gv.g = geom;
                    const NnResult<DIM, uint16_t, uint16_t> r = grid_nn_search_auto(gv, q.x, q.y, pz, seed);
                    d.x = gv.x[r.pos]; d.y = gv.y[r.pos];
                    idx = r.idx;
                    prev_pos[i] = (uint32_t)r.pos;
                }
                if (SHARED_TARGET) d2v[i] = d;
                hpart += mix64(((uint64_t)(uint32_t)i << 32) | (uint64_t)(uint32_t)idx);
                if (p.out_corr) {
                    if (p.corr_mode == 2) p.out_corr[corr_base + (uint64_t)it * (uint64_t)n + (uint64_t)i] = idx;
                    else if (p.corr_mode == 1 && (it + 1 == p.max_iter || p.fixed_point)) p.out_corr[corr_base + (uint64_t)i] = idx;
                }
            }
            hpart = warp_sum_u64(hpart);
            if (lane == 0) ctl->hash_part[warp] = hpart;
            __syncthreads();
            if (tid == 0) {
                uint64_t h = 0;
                for (int w = 0; w < NW; ++w) h += ctl->hash_part[w];
                ctl->H = ctl->H * 0x100000001B3ull + h;
                ctl->Hlast = h;
            }
            // the transformed source points of the thread's own slots (the search walked the scan with
            // every thread; the rows below belong to the workers)
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                double2 q;
                q.x = 0.; q.y = 0.;
                if (worker) q = r2v[wtid + k * WTHREADS];   // (service warps own no points and must not touch the rows)
                sx[k] = q.x; sy[k] = q.y;
            }   // (a worker reads and later overwrites only its own slots of r2v: no barrier needed here)
            ICP_LT_MARK(1)

// @region lean_irls
            // ---- estimate_transform (lib.rs:59-84)
            Transform Ti = transform_identity();
            double prev_error = kF64Max;
            uint32_t accepted = 0, why = kBreakCap;
            if (n < 2) {
                why = kBreakNone;  // check_input_size (lib.rs:186-189,225-228)
            } else {
                for (int inner = 0; inner < kInnerMaxIter; ++inner) {
// @region lean_stage0
                    // residuals (lib.rs:230-234): registers + shared rows (read by the service warps)
                    // fast: both bin maps usable (uniform: written before the previous evaluation's last barrier)
                    const bool fast = (sel->valid[0] & sel->valid[1]) != 0;
                    if (worker) {
                        const float lo0 = sel->lo0[0], lo1 = sel->lo0[1], s0 = sel->sc0[0], s1 = sel->sc0[1];
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            const int i = wtid + k * WTHREADS;
                            double tx_, ty_;
                            transform_apply(Ti, sx[k], sy[k], &tx_, &ty_);
                            const double2 d = matched(i);
                            double2 rr;
                            rr.x = tx_ - d.x;
                            rr.y = ty_ - d.y;
                            if (i < n) {
                                r2v[i] = rr;   // the accumulation below reads it back: the rows need no registers
                                if (fast) {
                                    const uint32_t b0 = lean_bin(rr.x, lo0, s0);
                                    const uint32_t b1 = lean_bin(rr.y, lo1, s1);
                                    lean_insert(h0, sel->bucket[0], b0, 1u, 0, i);
                                    lean_insert(h0, sel->bucket[1], b1, 0x10000u, 16, i);
                                    binv[i] = b0 | (b1 << 16);
                                }
                            }
                        }
                    }
                    if (fast) {
                        // the service warps wait for the whole histogram; everybody else goes on
                        if (warp < 2) icp_bar_sync(1, THREADS);
                        else icp_bar_arrive(1, THREADS);
                    }
                    ICP_LT_MARK(2)
                    if (fast && warp < 2) {
                        double v = 0., dev = 0.;
                        const float mlo = sel->lo0[warp], msc = sel->sc0[warp];
                        const bool ok = lean_serve(h0, sel->bucket[warp], binv, r2v, sel->s.list[warp], n, warp, 0, 0., &v,
                                                   sel->s.hist[warp]);
                        const bool ok1 = ok && lean_serve_mad(sel->bucket[warp], r2v, sel->s.list[warp], sel->s.hist[warp], n,
                                                              warp, v, mlo, msc, &dev);
                        if (lane == 0) {
                            sel->med[warp] = v; sel->ok[warp] = ok ? 1 : 0; sel->ok1[warp] = ok1 ? 1 : 0;
                            if (ok1) lean_publish(sel, warp, v, kPpf34 * dev);   // stats.rs:39-47
                        }
                    }
// @region lean_accumulate
                    // Huber weights (lib.rs:250), Jacobian (lib.rs:176-184), Huber cost (lib.rs:47-48):
                    // factorised partial sums, in the shadow of the service warps
                    double acc[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) acc[q] = 0.;
                    if (worker) {
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            const int i = wtid + k * WTHREADS;
                            lean_accumulate_point(acc, Ti, sx[k], sy[k], r2v[i], i < n);   // slots >= n hold zeros
                        }
                    }
                    warp_reduce16(acc);
                    if ((lane & 1) == 0) ctl->red[warp][warp_reduce16_owner(lane)] = acc[0];
                    __syncthreads();
                    ICP_LT_MARK(3)
                    bool served = false;
                    if (fast) {
                        if (sel->ok[0] & sel->ok[1] & sel->ok1[0] & sel->ok1[1]) {
                            served = true;
                            lean_zero_lh<THREADS>(sel->lh[0]);
                        } else if (sel->ok[0] & sel->ok[1]) {
// @region lean_stage1
                            const double m0 = sel->med[0], m1 = sel->med[1];
                            const float l0 = sel->lo1[0], l1 = sel->lo1[1], s0 = sel->sc1[0], s1 = sel->sc1[1];
                            lean_zero_lh<THREADS>(sel->lh[0]);
#pragma unroll
                            for (int k = 0; k < PPT; ++k) {
                                const int i = wtid + k * WTHREADS;
                                if (worker && i < n) {
                                    const double2 rr = r2v[i];
                                    const uint32_t b0 = lean_bin(fabs(rr.x - m0), l0, s0);
                                    const uint32_t b1 = lean_bin(fabs(rr.y - m1), l1, s1);
                                    lean_insert(h1, sel->bucket[0], b0, 1u, 0, i);
                                    lean_insert(h1, sel->bucket[1], b1, 0x10000u, 16, i);
                                    binv[i] = b0 | (b1 << 16);
                                }
                            }
                            __syncthreads();
                            ICP_LT_MARK(4)
                            if (warp < 2) {
                                double v = 0.;
                                const bool ok = lean_serve(h1, sel->bucket[warp], binv, r2v, sel->s.list[warp], n, warp, 1, warp ? m1 : m0, &v);
                                if (lane == 0) {
                                    sel->ok[warp] = ok ? 1 : 0;
                                    if (ok) lean_publish(sel, warp, warp ? m1 : m0, kPpf34 * v);   // stats.rs:39-47
                                }
                            }
                            __syncthreads();
                            ICP_LT_MARK(5)
                            served = (sel->ok[0] & sel->ok[1]) != 0;
                        }
                    }
                    if (!served) {
// @region lean_general_selection
                        lean_general_selection<THREADS, PPT>(sel, r2v, n);
                    } else if (!(sel->ok1[0] & sel->ok1[1])) {
                        lean_zero_lh<THREADS>(sel->lh[1]);
                    }
// @region lean_reduce
                    ICP_LT_MARK(6)
#ifdef ICP_LEAN_TIMING
                    ++lt_e;
#endif
                    lean_combine_solve<NW>(ctl, sel, Ti, prev_error, p.guard);
                    ICP_LT_MARK(7)
                    const uint32_t decision = (uint32_t)ctl->bcast[7];
                    if (decision != kBreakCap) { why = decision; break; }
                    prev_error = ctl->bcast[6];
                    Ti.r00 = ctl->bcast[0]; Ti.r01 = ctl->bcast[1]; Ti.r10 = ctl->bcast[2];
                    Ti.r11 = ctl->bcast[3]; Ti.tx = ctl->bcast[4];  Ti.ty = ctl->bcast[5];
                    ++accepted;
                }
            }
// @region lean_outer_tail
            const bool repeats = fused_repeats(p, accepted, it);   // icp_fused.cuh: the remaining iterations are copies
            if (p.out_cost && (it + 1 == p.max_iter || repeats)) {
                // huber_error (lib.rs:45-50) of the inner solve's final transform over this (the last)
                // iteration's transformed source / matched target points
                __syncthreads();   // every thread has read the last solve's broadcast
                double e = 0.;
                if (worker) {
#pragma unroll
                    for (int k = 0; k < PPT; ++k) {
                        const int i = wtid + k * WTHREADS;
                        double tx_, ty_;
                        transform_apply(Ti, sx[k], sy[k], &tx_, &ty_);
                        const double2 d = matched(i);
                        const double rx = tx_ - d.x, ry = ty_ - d.y;
                        const double rh = huber_rho(rx * rx + ry * ry);
                        e = e + (i < n ? rh : 0.);
                    }
                }
                const double tot = cta_tree_sum<NW>(ctl, e);
                if (tid == 0) p.out_cost[pair] = tot;
            }
            if (tid == 0) {   // lib.rs:127
                Transform T;
                T.r00 = ctl->T6[0]; T.r01 = ctl->T6[1]; T.r10 = ctl->T6[2]; T.r11 = ctl->T6[3]; T.tx = ctl->T6[4]; T.ty = ctl->T6[5];
                T = transform_mul(Ti, T);
                ctl->T6[0] = T.r00; ctl->T6[1] = T.r01; ctl->T6[2] = T.r10; ctl->T6[3] = T.r11; ctl->T6[4] = T.tx; ctl->T6[5] = T.ty;
                if (p.out_inner) p.out_inner[(size_t)pair * p.max_iter + it] = accepted | (why << 16);
                if (repeats) ctl->H = fused_repeat_hash(ctl->H, ctl->Hlast, p.max_iter - 1u - it);
            }
            if (repeats) fused_repeat_outputs<THREADS>(p, pair, it, n, corr_base, accepted | (why << 16));
            __syncthreads();  // rows are rewritten by the next search; ctl->bcast by the next solve; T6 is read
            if (repeats) break;
        }
        if (tid == 0) {
#pragma unroll
            for (int q = 0; q < 6; ++q) p.out_T[(size_t)pair * 6 + q] = ctl->T6[q];
            if (p.out_status) p.out_status[pair] = kPairOk;
            if (p.out_hash) p.out_hash[pair] = ctl->H;
            if (p.out_cost && p.max_iter == 0) p.out_cost[pair] = 0.;
            fused_report_knife(p, ctl, pair);
        }
        __syncthreads();  // ctl->pair is rewritten at the top of the loop
        ICP_LT_MARK(8)
#ifdef ICP_LEAN_TIMING
        ++lt_n;
#endif
    }
    if (tid == 0) fused_launch_exit(p);
#ifdef ICP_LEAN_TIMING
    if (blockIdx.x == 0 && tid == 0 && lt_n > 0)
        printf("lean timing CTA0: pairs %lld evals %lld | cycles/pair: setup+build %lld search %lld tail %lld | cycles/eval: stage0 %lld serve0 %lld stage1 %lld serve1 %lld accumulate %lld reduce+solve %lld\n",
               lt_n, lt_e, lt_acc[0] / lt_n, lt_acc[1] / lt_n, lt_acc[8] / lt_n, lt_acc[2] / lt_e, lt_acc[3] / lt_e,
               lt_acc[4] / lt_e, lt_acc[5] / lt_e, lt_acc[6] / lt_e, lt_acc[7] / lt_e);
#endif
}

}  // namespace icp
