// icp_fused.cuh — the whole ICP loop of one scan pair inside one CTA.
//
// Device counterpart of Icp2d::estimate (src/lib.rs:105-130) / Icp3d::estimate
// (src/lib.rs:148-173) with everything they call: estimate_transform (lib.rs:59-84),
// weighted_gauss_newton_update (lib.rs:218-261), calc_stddevs (stats.rs:49-60), huber
// (huber.rs), inverse3x3 (linalg.rs), Transform::new / mul (transform.rs).
//
// One persistent CTA takes scan pairs from an atomic counter.  Per pair:
//   build    : uniform grid over the target, counting-sorted into shared memory and mirrored to
//              an L2-resident per-CTA slab (Icp2d::new, lib.rs:97-102)
//   per outer iteration (lib.rs:112 — no early exit):
//     stage   the sorted target + cell table into shared memory
//     search  each thread transforms its PPT source points and finds their exact 1-NN
//             (points stay in registers: sx, sy, dx, dy)
//     IRLS    <=200 evaluations, each: residuals -> exact median/MAD per axis -> Huber row
//             weights -> 12 normal-equation sums + Huber cost, accumulated SEQUENTIALLY in
//             source order (i ascending, row 0 then row 1) by dedicated lanes so that every
//             sum carries the same roundings as the reference's fold -> 3x3 solve, break
//             tests, SE(2) exponential, compose.  No host round trip.
// Two reduction modes (template parameter ORDERED):
//   ORDERED = false  the normal-equation sums are reduced with warp shuffles in a fixed,
//             schedule-independent order (per-thread partials over the thread's points in
//             ascending order, xor-butterfly across lanes, warps added in warp order).  Weights,
//             Jacobian rows and terms never leave registers; the sorted target stays resident in
//             shared memory for the whole pair.  Results differ from the reference's sequential
//             fold only by summation rounding (~1e-16 relative): iteration counts and
//             correspondences match unless a break test falls within that distance of its
//             threshold, transforms agree to far better than 1e-9.
//   ORDERED = true   every sum carries the same roundings as the reference's fold: transforms are
//             bit-identical to the oracle by construction.  Shared memory is one region used as
//             {sorted target, cell table} during search and as {r, c, b, rho} row arrays during
//             IRLS (the target is re-staged from an L2-resident per-CTA slab each outer iteration).
#pragma once
#include "icp_target_nn.cuh"

namespace icp {

constexpr int kChainUnroll = 8;  // rows arrays are zero-padded to a multiple of 2x this many points
constexpr int kMaxWarps = 16;   // widest fused / lean instance: 512 threads

enum PairStatus : int32_t {
    kPairOk = 0,
    kPairEmptyTarget = 1,   // reference panics: index.unwrap() on an empty tree (lib.rs:122,165)
    kPairTooLarge = 2,      // pair routed to a kernel instance that cannot hold it
    kPairNonFinite = 3      // NaN / Inf input (reference panics in partial_cmp().unwrap(), stats.rs:12)
};

struct FusedLayout {
    // offsets from the start of the union region
    uint32_t nn_x, nn_y, nn_z, nn_idx, nn_cell, nn_bytes;
    uint32_t bt_cnt, bt_cell, bt_slot;
    uint32_t gn_r, gn_c, gn_b, gn_rho, gn_const;
    uint32_t u_bytes;
    // absolute offsets
    uint32_t off_sel, off_ctl, total;
};

struct FusedCtl {
    double sums[16];
    double red[kMaxWarps][16];
    double bcast[8];      // tree mode: inner transform after the update (6), cost (1), decision (1)
    uint64_t hash_part[kMaxWarps];
    uint32_t warp_tot[kMaxWarps];
    uint32_t pair;
    uint32_t knife;       // tree mode: a decision of this pair fell inside a guard band (GuardBands)
    double T6[6];         // lean schedule: the outer transform (kept out of the threads' registers)
    uint64_t H;           // lean schedule: running correspondence checksum
    uint64_t Hlast;       // lean schedule: the last outer iteration's term of the checksum
    uint32_t heavy_n;     // streaming kernel: source points left to a whole warp's search
};

// Knife-edge guard of the warp-tree mode.  The tree-reduced sums differ from the reference's
// sequential fold by rounding (~1e-16 relative per sum, up to ~1e-13 on a decision quantity once the
// perturbation has travelled through a few updates), and the inner loop's decisions are
// discontinuous: det == 0 (linalg.rs:18), delta.dot(delta) < 1e-6 (lib.rs:71), error > prev_error
// (lib.rs:76).  Whenever a decision quantity lies within a relative band of its threshold — bands
// three or more orders of magnitude wider than that perturbation — the pair is reported in
// FusedParams::redo_list and the host re-runs it with the reference-order sums (ORDERED), whose
// decisions and transforms are the reference's by construction.  A pair that is never flagged took
// every decision with a margin the summation order cannot bridge.  0 disables a band.
struct GuardBands {
    double det;     // |det| <= det * (|m00 n00| + |m10 c01| + |m20 n02|)
    double delta;   // |d.d - 1e-6| <= delta * 1e-6
    double err;     // |error - prev_error| <= err * max(error, prev_error)
};
// The library's bands.  Measured drift of the tree-mode state against the reference-order state
// over 131,072 config-2 pairs (tools/census.py, profiles/): transforms differ by <= 6.3e-13 and
// the final huber_error by ~1e-13 relative; the bands sit two orders of magnitude above that.
constexpr double kGuardDet = 1e-10, kGuardDelta = 1e-8, kGuardErr = 1e-11;

ICP_HD uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }
ICP_HD uint32_t align128(uint32_t v) { return (v + 127u) & ~127u; }

ICP_HD FusedLayout fused_layout(int dim, int ncap, int mcap, int cell_budget, bool shared_target, bool ordered)
{
    FusedLayout L;
    uint32_t o = 0;
    if (shared_target) { mcap = 0; cell_budget = 0; }
    L.nn_x = o; o = align16(o + 8u * (uint32_t)mcap);
    L.nn_y = o; o = align16(o + 8u * (uint32_t)mcap);
    L.nn_z = o; if (dim == 3) o = align16(o + 8u * (uint32_t)mcap);
    L.nn_idx = o; o = align16(o + 2u * (uint32_t)mcap);
    L.nn_cell = o; o = align16(o + 2u * (uint32_t)(cell_budget + 1));
    L.nn_bytes = o;
    o = align128(o);
    const uint32_t bt0 = o;
    L.bt_cnt = o; o = align16(o + 4u * (uint32_t)cell_budget);
    L.bt_cell = o; o = align16(o + 2u * (uint32_t)mcap);
    L.bt_slot = o; o = align16(o + 2u * (uint32_t)mcap);
    const uint32_t nn_total = o;
    // IRLS row arrays, one double2 (row 0, row 1) per point.
    //  ordered: r, b, c, rho and the constant table overlay the whole search region; the chain
    //           warp reads them with 128-bit loads whose lanes address different arrays at the
    //           same point index, so the arrays sit at different 16-byte phases of the 128-byte
    //           bank line: r +0, b +16, K0 +32, K1 +48 (c is read alone).
    //  tree:    only r is kept, overlaying the build temporaries; the search arrays stay resident.
    o = ordered ? 0u : bt0;
    L.gn_r = o; o = align128(o + 16u * (uint32_t)ncap);
    if (ordered) {
        L.gn_b = o + 16u; o = align128(L.gn_b + 16u * (uint32_t)ncap);
        L.gn_c = o; o = align128(o + 16u * (uint32_t)ncap);
        L.gn_rho = o; o = align128(o + 8u * (uint32_t)ncap);
        L.gn_const = o + 32u; o = align128(L.gn_const + 16u + 2u * 16u * (uint32_t)kChainUnroll);
    } else {
        L.gn_b = L.gn_c = L.gn_rho = L.gn_const = o;
    }
    L.u_bytes = o > nn_total ? o : nn_total;
    L.off_sel = L.u_bytes;
    L.off_ctl = align16(L.off_sel + (uint32_t)sizeof(SelShared));
    L.total = align16(L.off_ctl + (uint32_t)sizeof(FusedCtl));
    return L;
}

struct FusedParams {
    const double* src;          // AoS, DIM doubles per point
    const uint32_t* src_off;    // [B+1] in points
    const double* dst;          // AoS (ignored with a shared target)
    const uint32_t* dst_off;    // [B+1]
    const double* init_T;       // [B*6]  r00 r01 r10 r11 tx ty
    double* out_T;              // [B*6]
    uint32_t* out_inner;        // [B*max_iter] accepted | reason << 16 ; nullable
    int32_t* out_corr;          // nullable
    uint64_t* out_hash;         // [B] nullable
    int32_t* out_status;        // [B] nullable
    int32_t* out_error;         // one word: max pair status of the launch (zeroed by the host)
    uint32_t* work_counter;     // zeroed by the host before launch
    unsigned char* workspace;   // gridDim.x * ws_stride bytes (per-CTA sorted-target slab)
    uint64_t ws_stride;
    TargetView target;          // shared target
    uint32_t pair_begin, pair_end;   // this launch handles pairs [pair_begin, pair_end)
    uint32_t max_iter;
    int corr_mode;              // 0 none, 1 last iteration, 2 all iterations
    int cell_budget;
    int ncap, mcap;             // capacities the layout was built for
    uint32_t sm_count;          // SMs of the device (residency-slot estimate for warp-role rotation)
    // pair_list != null: the launch handles pairs pair_list[0 .. *pair_count) instead of
    // [pair_begin, pair_end) — the ordered re-run of the pairs a tree-mode launch flagged
    const uint32_t* pair_list;
    const uint32_t* pair_count;
    // tree mode: pairs with a decision inside a guard band are appended here (null: no guard)
    uint32_t* redo_list;
    uint32_t* redo_count;
    GuardBands guard;
    double* out_cost;           // [B] nullable: huber_error (lib.rs:45-50) of the last inner solve
    uint32_t shared_src;        // every pair aligns the scan src[0 .. src_off[1])
    // Self-cleaning counters (device-pointer calls, which run back to back without a host memset):
    // the last CTA to leave the launch zeroes work_counter and fin_counter; the re-run launch also
    // moves *pair_count to *redo_seen and zeroes it.  Null: the host zeroes the counters per call.
    uint32_t* fin_counter;
    uint32_t* redo_seen;
    uint32_t fixed_point;       // 1: outer iterations that provably repeat are written, not recomputed (below)
};

// Fixed point of the outer loop (exact, not a convergence test).  An inner solve that accepts no
// update (lib.rs:67-77 breaks in its first pass) returns the identity, so lib.rs:127 leaves the
// outer transform as it was: the next outer iteration transforms the scan to the same points,
// finds the same correspondences, runs the same inner solve and ends the same way — and so does
// every iteration after it.  The reference has no exit and computes them all (lib.rs:112-128);
// the kernels write what they would produce — the same correspondences and inner-iteration word
// per remaining iteration, the checksum advanced once per iteration — and go to the next pair.
// `transform_mul(identity, T)` is still applied once, as the reference does (it turns a -0.0 of a
// caller's initial transform into +0.0; a second application changes nothing).
// ICP_B200_FIXED_POINT=0 computes every iteration.
ICP_DEV bool fused_repeats(const FusedParams& p, uint32_t accepted, uint32_t it)
{
    return p.fixed_point != 0u && accepted == 0u && it + 1u < p.max_iter;
}

// Outputs of outer iterations it+1 .. max_iter-1, copies of iteration `it` (whole CTA; the
// correspondences of iteration `it` were written before a barrier this call sits behind).
template <int THREADS>
ICP_DEV void fused_repeat_outputs(const FusedParams& p, uint32_t pair, uint32_t it, int n, uint64_t corr_base, uint32_t word)
{
    const int tid = (int)threadIdx.x;
    if (p.out_corr && p.corr_mode == 2) {
        const int32_t* from = p.out_corr + corr_base + (uint64_t)it * (uint64_t)n;
        for (uint32_t j = it + 1u; j < p.max_iter; ++j) {
            int32_t* to = p.out_corr + corr_base + (uint64_t)j * (uint64_t)n;
            for (int i = tid; i < n; i += THREADS) to[i] = from[i];
        }
    }
    if (tid == 0 && p.out_inner)
        for (uint32_t j = it + 1u; j < p.max_iter; ++j) p.out_inner[(size_t)pair * p.max_iter + j] = word;
}

// The checksum after `times` more iterations that each contribute the term h.
ICP_DEV uint64_t fused_repeat_hash(uint64_t H, uint64_t h, uint32_t times)
{
    for (uint32_t j = 0; j < times; ++j) H = H * 0x100000001B3ull + h;
    return H;
}

// Entry of every ICP launch.  With programmatic dependent launch (ICP_B200_FLAG_INDEPENDENT) the
// next call's CTAs fill the SMs this call's last pairs leave idle.  Two rules keep the library's
// per-call state (two alternating sets of counters, workspaces and flagged-pair lists) safe:
//  - the re-run launch reads the list its main launch wrote, so all of it waits for that launch;
//  - CTA 0 of a main launch waits for the previous call to complete BEFORE it releases its own
//    dependents, so the call after this one cannot start while the call before it still runs.
ICP_DEV void fused_launch_entry(const FusedParams& p)
{
    if (p.pair_list) {
        icp_pdl_launch_dependents();
        icp_pdl_wait();
    } else {
        if (blockIdx.x == 0) icp_pdl_wait();
        icp_pdl_launch_dependents();
    }
}

// Exit of every ICP launch (thread 0 of each CTA, after its last pair).
ICP_DEV void fused_launch_exit(const FusedParams& p)
{
    if (!p.fin_counter) return;
    __threadfence();
    if (atomicAdd(p.fin_counter, 1u) + 1u == gridDim.x) {
        *p.work_counter = 0u;
        *p.fin_counter = 0u;
        if (p.pair_list) {
            uint32_t* count = const_cast<uint32_t*>(p.pair_count);
            *p.redo_seen = *count;
            *count = 0u;
        }
        __threadfence();
    }
}

// Next pair of this launch (thread 0 only): a contiguous range, or the flagged-pair list.
// Returns a value >= p.pair_end when the work is exhausted.
ICP_DEV uint32_t fused_next_pair(const FusedParams& p)
{
    const uint32_t k = atomicAdd(p.work_counter, 1u);
    if (p.pair_list) return k < *p.pair_count ? p.pair_list[k] : 0xffffffffu;
    return p.pair_begin + k;
}

// End of a pair in tree mode (thread 0 only): report it for the ordered re-run if it was flagged.
ICP_DEV void fused_report_knife(const FusedParams& p, FusedCtl* ctl, uint32_t pair)
{
    if (ctl->knife && p.redo_list) p.redo_list[atomicAdd(p.redo_count, 1u)] = pair;
}

// ---------------------------------------------------------------------------------------------
// @region build_grid
template <int THREADS, int DIM>
ICP_DEV GridGeom cta_build_grid(const double* dst, int m, unsigned char* U, const FusedLayout& L,
                                FusedCtl* ctl, int cell_budget)
{
    constexpr int NW = THREADS / 32;
    const int tid = (int)threadIdx.x, lane = lane_id(), warp = warp_id();
    const double inf = 1.0 / 0.0;
    double mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
    for (int j = tid; j < m; j += THREADS) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double v = dst[(size_t)j * DIM + d];
            mn[d] = fmin(mn[d], v);
            mx[d] = fmax(mx[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        mn[d] = warp_min_f64(mn[d]);
        mx[d] = warp_max_f64(mx[d]);
    }
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            ctl->red[warp][d] = mn[d];
            ctl->red[warp][3 + d] = mx[d];
        }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 3; ++d) { mn[d] = inf; mx[d] = -inf; }
    for (int w = 0; w < NW; ++w) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            mn[d] = fmin(mn[d], ctl->red[w][d]);
            mx[d] = fmax(mx[d], ctl->red[w][3 + d]);
        }
    }
    const GridGeom g = grid_make_geom(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], DIM, cell_budget);
    const int ncells = g.gx * g.gy * g.gz;

    uint32_t* cnt = reinterpret_cast<uint32_t*>(U + L.bt_cnt);
    uint16_t* tcell = reinterpret_cast<uint16_t*>(U + L.bt_cell);
    uint16_t* tslot = reinterpret_cast<uint16_t*>(U + L.bt_slot);
    double* sx = reinterpret_cast<double*>(U + L.nn_x);
    double* sy = reinterpret_cast<double*>(U + L.nn_y);
    double* sz = reinterpret_cast<double*>(U + L.nn_z);
    uint16_t* sidx = reinterpret_cast<uint16_t*>(U + L.nn_idx);
    uint16_t* cstart = reinterpret_cast<uint16_t*>(U + L.nn_cell);

    for (int c = tid; c < ncells; c += THREADS) cnt[c] = 0u;
    __syncthreads();
    for (int j = tid; j < m; j += THREADS) {
        int cx = grid_coord(dst[(size_t)j * DIM + 0], g.ox, g.inv_h, g.gx);
        int cy = grid_coord(dst[(size_t)j * DIM + 1], g.oy, g.inv_h, g.gy);
        int cz = (DIM == 3) ? grid_coord(dst[(size_t)j * DIM + 2], g.oz, g.inv_h, g.gz) : 0;
        int cell = (cz * g.gy + cy) * g.gx + cx;
        uint32_t slot = atomicAdd(&cnt[cell], 1u);
        tcell[j] = (uint16_t)cell;
        tslot[j] = (uint16_t)slot;
    }
    __syncthreads();
    // exclusive scan of cnt -> cstart, one contiguous chunk per warp
    const int chunk = (((ncells + NW - 1) / NW) + 31) & ~31;
    const int c0 = warp * chunk;
    {
        uint32_t s = 0;
        for (int c = c0 + lane; c < c0 + chunk; c += 32) s += (c < ncells) ? cnt[c] : 0u;
        s = warp_sum_u32(s);
        if (lane == 0) ctl->warp_tot[warp] = s;
    }
    __syncthreads();
    {
        uint32_t base = 0;
        for (int w = 0; w < warp; ++w) base += ctl->warp_tot[w];
        for (int c = c0 + lane; c < c0 + chunk; c += 32) {
            uint32_t v = (c < ncells) ? cnt[c] : 0u;
            uint32_t incl = warp_inclusive_scan_u32(v);
            if (c < ncells) cstart[c] = (uint16_t)(base + incl - v);
            base += __shfl_sync(ICP_FULL_MASK, incl, 31);
        }
        if (tid == 0) cstart[ncells] = (uint16_t)m;
    }
    __syncthreads();
    for (int j = tid; j < m; j += THREADS) {
        int pos = (int)cstart[tcell[j]] + (int)tslot[j];
        sx[pos] = dst[(size_t)j * DIM + 0];
        sy[pos] = dst[(size_t)j * DIM + 1];
        if (DIM == 3) sz[pos] = dst[(size_t)j * DIM + 2];
        sidx[pos] = (uint16_t)j;
    }
    __syncthreads();
    return g;
}

// ---------------------------------------------------------------------------------------------
// Sequential accumulation (lib.rs:242-255 and lib.rs:45-50).  Lane L < 12 of the chain warp owns
// one accumulator (9 of jtj, 3 of jtr) and adds its terms in the reference's order.  Term of
// row (i, j):
//     (X * Y) * Z   with  X = c_ij = w_ij * g_j,   Y = J_ij[a],   Z = J_ij[b] or r_ij
// J_ij = [R_j0, R_j1, b_ij].  Rows 0 and 1 of a point are one double2, so each factor is one
// 128-bit shared load per point; uniform factors (R_j0, R_j1) come from a small table whose
// entries repeat per unrolled point, so that every lane executes the same loads with the same
// immediates and only the base pointer differs.  Lanes >= 12 mirror lane 0 (same addresses,
// result discarded).
// @region chain
struct ChainPtrs {
    smem_ref x, y, z;   // shared-memory references to double2 arrays
    int iy, iz;         // advance (in double2) per kChainStage points; x always advances
};

constexpr int kChainStage = 4;   // points per pipeline stage

// Three-stage software pipeline over stages of kChainStage points: while stage k's dependent
// additions retire, stage k+1's (independent) products are formed and stage k+2's operands are
// loaded, so the loop runs at the latency of the DADD chain alone.  Kept out of line so that it
// gets its own register allocation instead of competing with the per-point state the calling
// kernel keeps in registers.  nstages >= 2.
ICP_DEV_NOINLINE double chain_run(ChainPtrs p, int nstages, double acc = 0.0)
{
    constexpr int S = kChainStage;
    double2 lx[S], ly[S], lz[S];
    double t[2 * S];
#pragma unroll
    for (int u = 0; u < S; ++u) {
        const double2 x = smem_ld2(p.x, u), y = smem_ld2(p.y, u), z = smem_ld2(p.z, u);
        t[2 * u] = (x.x * y.x) * z.x;
        t[2 * u + 1] = (x.y * y.y) * z.y;
    }
    p.x = smem_adv2(p.x, S); p.y = smem_adv2(p.y, p.iy); p.z = smem_adv2(p.z, p.iz);
#pragma unroll
    for (int u = 0; u < S; ++u) { lx[u] = smem_ld2(p.x, u); ly[u] = smem_ld2(p.y, u); lz[u] = smem_ld2(p.z, u); }
    for (int st = 2; st < nstages; ++st) {
        p.x = smem_adv2(p.x, S); p.y = smem_adv2(p.y, p.iy); p.z = smem_adv2(p.z, p.iz);
        double tn[2 * S];
#pragma unroll
        for (int u = 0; u < S; ++u) {
            tn[2 * u] = (lx[u].x * ly[u].x) * lz[u].x;
            tn[2 * u + 1] = (lx[u].y * ly[u].y) * lz[u].y;
            lx[u] = smem_ld2(p.x, u); ly[u] = smem_ld2(p.y, u); lz[u] = smem_ld2(p.z, u);
            acc = acc + t[2 * u];
            acc = acc + t[2 * u + 1];
        }
#pragma unroll
        for (int u = 0; u < 2 * S; ++u) t[u] = tn[u];
    }
#pragma unroll
    for (int u = 0; u < 2 * S; ++u) acc = acc + t[u];
#pragma unroll
    for (int u = 0; u < S; ++u) {
        acc = acc + (lx[u].x * ly[u].x) * lz[u].x;
        acc = acc + (lx[u].y * ly[u].y) * lz[u].y;
    }
    return acc;
}

// sum of the Huber cost terms in source order (lib.rs:45-50); n2 = number of double2 entries
// (a multiple of kChainUnroll)
ICP_DEV_NOINLINE double cost_run(smem_ref rp, int n2, double e = 0.0)
{
    double2 v[kChainUnroll];
#pragma unroll
    for (int u = 0; u < kChainUnroll; ++u) v[u] = smem_ld2(rp, u);
    for (int i = kChainUnroll; i < n2; i += kChainUnroll) {
        rp = smem_adv2(rp, kChainUnroll);
#pragma unroll
        for (int u = 0; u < kChainUnroll; ++u) {
            const double2 w = v[u];
            v[u] = smem_ld2(rp, u);
            e = e + w.x;
            e = e + w.y;
        }
    }
#pragma unroll
    for (int u = 0; u < kChainUnroll; ++u) {
        e = e + v[u].x;
        e = e + v[u].y;
    }
    return e;
}

// ktab: K0[kChainStage] = (R00, R10) repeated, then (after a 16-byte phase shift) K1 = (R01, R11)
constexpr int kChainK1 = kChainUnroll + 1;   // offset of K1 in double2 units

ICP_DEV ChainPtrs chain_role(int lane, const double2* r2, const double2* c2, const double2* b2,
                             const double2* ktab)
{
    ChainPtrs p;
    const int l = lane < 12 ? lane : 0;
    const int a = (l < 9) ? l / 3 : l - 9;
    const int b = (l < 9) ? l % 3 : 3;
    p.x = smem_ref_of(c2);
    if (a < 2) { p.y = smem_ref_of(ktab + a * kChainK1); p.iy = 0; } else { p.y = smem_ref_of(b2); p.iy = kChainStage; }
    if (b < 2) { p.z = smem_ref_of(ktab + b * kChainK1); p.iz = 0; }
    else { p.z = smem_ref_of((b == 2) ? b2 : r2); p.iz = kChainStage; }
    return p;
}

// ---------------------------------------------------------------------------------------------
// Warp-tree mode building blocks (shared by the fused and the streaming kernel).
//
// One point's contribution to the sums of lib.rs:242-255 and lib.rs:45-50.  acc: 0-5 upper
// triangle of J^T W J (00 01 02 11 12 22), 6-8 J^T W r, 9 Huber cost, 10-15 unused (zero).
// J^T W J is accumulated as its upper triangle and mirrored: (c J_a) J_b and (c J_b) J_a differ by
// one rounding per term, the same order as the tree reduction's own reordering.
// @region tree_accumulate
// `valid` false (a thread's slot beyond the scan's last point; its coordinates are zeros, its
// residual is finite): every term is formed and enters as c = 0, i.e. as +-0, which leaves the
// sums unchanged.  No branch on it, so the PPT points of a thread form one basic block and their
// quotients and square roots overlap.
ICP_DEV void tree_accumulate_point(double acc[16], const Transform& Ti, double sx, double sy, double2 rr,
                                   const double sigma[2], double g0, double g1, bool valid = true)
{
    const double a0 = -sy, a1 = sx;   // jacobian(), lib.rs:176-184
    const double J0[3] = {Ti.r00, Ti.r01, Ti.r00 * a0 + Ti.r01 * a1};
    const double J1[3] = {Ti.r10, Ti.r11, Ti.r10 * a0 + Ti.r11 * a1};
    if (sigma[0] != 0.) {
        const double cw = huber_drho_of_residual(rr.x) * g0;
        const double c = valid ? cw : 0.;
        const double p0 = c * J0[0], p1 = c * J0[1], p2 = c * J0[2];
        acc[0] = acc[0] + p0 * J0[0]; acc[1] = acc[1] + p0 * J0[1]; acc[2] = acc[2] + p0 * J0[2];
        acc[3] = acc[3] + p1 * J0[1]; acc[4] = acc[4] + p1 * J0[2]; acc[5] = acc[5] + p2 * J0[2];
        acc[6] = acc[6] + p0 * rr.x; acc[7] = acc[7] + p1 * rr.x; acc[8] = acc[8] + p2 * rr.x;
    }
    if (sigma[1] != 0.) {
        const double cw = huber_drho_of_residual(rr.y) * g1;
        const double c = valid ? cw : 0.;
        const double p0 = c * J1[0], p1 = c * J1[1], p2 = c * J1[2];
        acc[0] = acc[0] + p0 * J1[0]; acc[1] = acc[1] + p0 * J1[1]; acc[2] = acc[2] + p0 * J1[2];
        acc[3] = acc[3] + p1 * J1[1]; acc[4] = acc[4] + p1 * J1[2]; acc[5] = acc[5] + p2 * J1[2];
        acc[6] = acc[6] + p0 * rr.y; acc[7] = acc[7] + p1 * rr.y; acc[8] = acc[8] + p2 * rr.y;
    }
    const double rho = huber_rho(rr.x * rr.x + rr.y * rr.y);
    acc[9] = acc[9] + (valid ? rho : 0.);
}

// @region solve3x3
// The same solve by one full warp whose lanes hold identical arguments: all lanes form the nine
// cofactor numerators, lane q < 9 forms the one quotient i_q = numerator_q / det and the warp
// exchanges them — one division on the critical path instead of nine.  Same operations on the
// same operands as solve_normal_equations, hence the same bits.
ICP_DEV bool solve_normal_equations_warp(const double m[9], const double jtr[3], double delta[3],
                                         double* det_out = nullptr, double* det_scale = nullptr)
{
    const double m00 = m[0], m01 = m[1], m02 = m[2];
    const double m10 = m[3], m11 = m[4], m12 = m[5];
    const double m20 = m[6], m21 = m[7], m22 = m[8];
    const double n00 = m22 * m11 - m21 * m12;
    const double c01 = m22 * m01 - m21 * m02;
    const double n02 = m12 * m01 - m11 * m02;
    const double det = m00 * n00 - m10 * c01 + m20 * n02;
    if (det_out) {
        *det_out = det;
        *det_scale = (fabs(m00 * n00) + fabs(m10 * c01)) + fabs(m20 * n02);
    }
    if (det == 0.) return false;
    const double n01 = -c01;
    const double n10 = -(m22 * m10 - m20 * m12);
    const double n11 = m22 * m00 - m20 * m02;
    const double n12 = -(m12 * m00 - m10 * m02);
    const double n20 = m21 * m10 - m20 * m11;
    const double n21 = -(m21 * m00 - m20 * m01);
    const double n22 = m11 * m00 - m10 * m01;
    const int lane = lane_id();
    double num = n00;
    num = lane == 1 ? n01 : num; num = lane == 2 ? n02 : num;
    num = lane == 3 ? n10 : num; num = lane == 4 ? n11 : num; num = lane == 5 ? n12 : num;
    num = lane == 6 ? n20 : num; num = lane == 7 ? n21 : num; num = lane == 8 ? n22 : num;
    const double iq = num / det;
    const double i00 = shfl_f64(iq, 0), i01 = shfl_f64(iq, 1), i02 = shfl_f64(iq, 2);
    const double i10 = shfl_f64(iq, 3), i11 = shfl_f64(iq, 4), i12 = shfl_f64(iq, 5);
    const double i20 = shfl_f64(iq, 6), i21 = shfl_f64(iq, 7), i22 = shfl_f64(iq, 8);
    delta[0] = ((-i00) * jtr[0] + (-i01) * jtr[1]) + (-i02) * jtr[2];
    delta[1] = ((-i10) * jtr[0] + (-i11) * jtr[1]) + (-i12) * jtr[2];
    delta[2] = ((-i20) * jtr[0] + (-i21) * jtr[1]) + (-i22) * jtr[2];
    return true;
}

// The decisions of one inner iteration (lib.rs:67-81) by one full warp whose lanes hold identical
// sums: solve, the two break tests with their guard bands, the update; lane 0 publishes
// {new inner transform, cost, decision} in ctl->bcast and raises ctl->knife.
ICP_DEV void tree_decide(FusedCtl* ctl, const double jtj[9], const double jtr[3], double error, const Transform& Ti,
                         double prev_error, const GuardBands& gb)
{
    const int lane = lane_id();
    double delta[3];
    uint32_t decision = kBreakCap;   // kBreakCap here = "update accepted"
    Transform Tn = Ti;
    double det, det_scale;
    bool knife;
    if (!solve_normal_equations_warp(jtj, jtr, delta, &det, &det_scale)) {
        decision = kBreakNone;
        knife = det_scale > 0. && gb.det > 0.;   // all-zero sums (sigma == 0 rows, lib.rs:245-247) are zero in any order
    } else {
        const double dd = (delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2];
        knife = fabs(det) <= gb.det * det_scale || fabs(dd - kDeltaNormThreshold) <= gb.delta * kDeltaNormThreshold;
        if (dd < kDeltaNormThreshold) {
            decision = kBreakDelta;
        } else {
            knife = knife || fabs(error - prev_error) <= gb.err * fmax(error, prev_error);
            if (error > prev_error) decision = kBreakError;
            else Tn = transform_mul(transform_new(delta[0], delta[1], delta[2]), Ti);
        }
    }
    if (lane == 0 && knife) ctl->knife = 1u;
    if (lane == 0) {
        ctl->bcast[0] = Tn.r00; ctl->bcast[1] = Tn.r01; ctl->bcast[2] = Tn.r10;
        ctl->bcast[3] = Tn.r11; ctl->bcast[4] = Tn.tx;  ctl->bcast[5] = Tn.ty;
        ctl->bcast[6] = error;
        ctl->bcast[7] = (double)decision;
    }
}

// Reduces the threads' partial sums in a fixed tree (split butterfly across lanes, warps in warp
// order), then warp 0 solves and takes the break decisions of lib.rs:67-81 once and publishes
// {new inner transform, cost, decision} in ctl->bcast.  Two barriers.  Callers must not write
// ctl->red / ctl->bcast between two calls without an intervening barrier of their own (the
// selection of the next evaluation provides several).
// @region tree_reduce
template <int NW>
ICP_DEV void cta_tree_reduce_solve(FusedCtl* ctl, double acc[16], const Transform& Ti, double prev_error,
                                   const GuardBands& gb)
{
    const int lane = lane_id(), warp = warp_id();
    warp_reduce16(acc);
    if ((lane & 1) == 0) ctl->red[warp][warp_reduce16_owner(lane)] = acc[0];
    __syncthreads();
    if (warp == 0) {
        double tot = 0.;
        if (lane < 16) {
            tot = ctl->red[0][lane];
            for (int w = 1; w < NW; ++w) tot = tot + ctl->red[w][lane];
        }
        double v[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) v[q] = shfl_f64(tot, q);
        const double jtj[9] = {v[0], v[1], v[2], v[1], v[3], v[4], v[2], v[4], v[5]};
        const double jtr[3] = {v[6], v[7], v[8]};
        const double error = v[9];
        tree_decide(ctl, jtj, jtr, error, Ti, prev_error, gb);
    }
    __syncthreads();
}

// Fixed-tree sum of one value per thread (xor butterfly across lanes, warps in warp order): the
// tree-mode form of the final huber_error.  Every thread returns the total.  Two barriers; uses
// ctl->red / ctl->bcast like cta_tree_reduce_solve.
template <int NW>
ICP_DEV double cta_tree_sum(FusedCtl* ctl, double v)
{
    const int lane = lane_id(), warp = warp_id();
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = v + shfl_xor_f64(v, m);
    if (lane == 0) ctl->red[warp][0] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = ctl->red[0][0];
        for (int w = 1; w < NW; ++w) tot = tot + ctl->red[w][0];
        ctl->bcast[6] = tot;
    }
    __syncthreads();
    return ctl->bcast[6];
}

// ---------------------------------------------------------------------------------------------
// calc_stddevs (stats.rs:49-60) over the CTA's residuals: per-axis median, then the median of
// |r - median|, sigma = 1.4826 * MAD (stats.rs:39-47).  est_med / est_sigma carry the previous
// evaluation's result in and this one's out: they only steer the (exact) selection's bin map.
// One inlined copy of the selection runs both stages.
// @region calc_stddevs_driver
template <int THREADS, int PPT, class KeySrc>
ICP_DEV int cta_calc_stddevs(SelShared* sel, const KeySrc& ks, int n, double est_med[2], double est_sigma[2],
                             int hist_par, double sigma[2])
{
    double med[2] = {0., 0.}, out[2] = {0., 0.};
#pragma unroll 1
    for (int stage = 0; stage < 2; ++stage) {
        bool ev[2];
        float elo[2], esc[2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
            ev[a] = stage == 0 ? sel_fast_params(est_med[a] - 3.0 * est_sigma[a], 6.0 * est_sigma[a], &elo[a], &esc[a])
                               : sel_fast_params(0.0, 3.0 * est_sigma[a], &elo[a], &esc[a]);
        hist_par = block_median2<THREADS, PPT>(sel, ks, n, stage, med, ev, elo, esc, hist_par, out);
        if (stage == 0) { med[0] = out[0]; med[1] = out[1]; }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        sigma[a] = kPpf34 * out[a];
        est_med[a] = med[a];
        est_sigma[a] = sigma[a];
    }
    return hist_par;
}

// ---------------------------------------------------------------------------------------------
// @region pair_setup
template <int THREADS, int PPT, int DIM, bool SHARED_TARGET, int MINB, bool ORDERED>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(THREADS, MINB) icp_fused_kernel(FusedParams p)
{
    constexpr int NW = THREADS / 32;
    static_assert(NW <= kMaxWarps, "FusedCtl holds per-warp slots for kMaxWarps warps");
    constexpr int NCAP = THREADS * PPT;
    static_assert(NCAP % (2 * kChainUnroll) == 0, "row arrays are padded to twice the chain unroll");
    ICP_SHARED_DYN(smem);
    const FusedLayout L = fused_layout(DIM, NCAP, p.mcap, p.cell_budget, SHARED_TARGET, ORDERED);
    unsigned char* U = smem;
    SelShared* sel = reinterpret_cast<SelShared*>(smem + L.off_sel);
    FusedCtl* ctl = reinterpret_cast<FusedCtl*>(smem + L.off_ctl);
    double* rho = reinterpret_cast<double*>(U + L.gn_rho);
    double2* r2v = reinterpret_cast<double2*>(U + L.gn_r);
    double2* c2v = reinterpret_cast<double2*>(U + L.gn_c);
    double2* b2v = reinterpret_cast<double2*>(U + L.gn_b);
    double2* ktab = reinterpret_cast<double2*>(U + L.gn_const);
    // chain / cost-sum warps rotate with the CTA's residency slot so that co-resident CTAs put
    // their serial phases on different SM sub-partitions (warp w issues on sub-partition w % 4)
    const int slot = (int)(blockIdx.x / (p.sm_count > 0 ? p.sm_count : 1u));
    const int chain_warp = slot % NW;
    const int cost_warp = (slot + 1) % NW;

    const int tid = (int)threadIdx.x, lane = lane_id(), warp = warp_id();
    unsigned char* slab = p.workspace + (size_t)blockIdx.x * p.ws_stride;

    for (int q = tid; q < 2 * kSelHistPad; q += THREADS) (&sel->hist[0][0])[q] = 0u;
    int hist_par = 0;
    fused_launch_entry(p);
    __syncthreads();

    for (;;) {
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
        Transform T;
        T.r00 = p.init_T[(size_t)pair * 6 + 0]; T.r01 = p.init_T[(size_t)pair * 6 + 1];
        T.r10 = p.init_T[(size_t)pair * 6 + 2]; T.r11 = p.init_T[(size_t)pair * 6 + 3];
        T.tx = p.init_T[(size_t)pair * 6 + 4];  T.ty = p.init_T[(size_t)pair * 6 + 5];

        int status = kPairOk;
        if (m == 0 && n > 0 && p.max_iter > 0) status = kPairEmptyTarget;
        if (n > NCAP || (!SHARED_TARGET && m > p.mcap)) status = kPairTooLarge;
        {
            // the reference panics on NaN residuals; reject non-finite input up front
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

        // ---- Icp2d::new / Icp3d::new: sorted cell list of the target
        GridGeom geom;
        if (!SHARED_TARGET) {
            if (m > 0) {
                geom = cta_build_grid<THREADS, DIM>(dst, m, U, L, ctl, p.cell_budget);
                if (ORDERED)
                    for (uint32_t o = (uint32_t)tid * 16u; o < L.nn_bytes; o += THREADS * 16u)
                        *reinterpret_cast<uint4*>(slab + o) = *reinterpret_cast<const uint4*>(U + o);
            }
        } else {
            geom = p.target.g;
        }

        double est_med[2] = {0., 0.}, est_sigma[2] = {0., 0.};
        uint64_t H = 0;  // meaningful in thread 0
        const int npad = ((n + 2 * kChainUnroll - 1) / (2 * kChainUnroll)) * (2 * kChainUnroll);
        const uint64_t corr_base = (p.shared_src ? (uint64_t)pair * (uint64_t)n : (uint64_t)p.src_off[pair]) *
                                   (p.corr_mode == 2 ? p.max_iter : 1u);

        int ppos[PPT];   // sorted position of each point's match in the previous outer iteration (ordered mode)
#pragma unroll
        for (int k = 0; k < PPT; ++k) ppos[k] = -1;
        for (uint32_t it = 0; it < p.max_iter; ++it) {
            // ---- stage the sorted target (the union region was overwritten by the IRLS rows)
            if (ORDERED && !SHARED_TARGET && it > 0 && m > 0) {
                for (uint32_t o = (uint32_t)tid * 16u; o < L.nn_bytes; o += THREADS * 16u)
                    *reinterpret_cast<uint4*>(U + o) = *reinterpret_cast<const uint4*>(slab + o);
                __syncthreads();
            }
// @region search
            // ---- transform (lib.rs:113-116) + correspondence search (lib.rs:118-124)
            double sx[PPT], sy[PPT], dx[PPT], dy[PPT];
            uint64_t hpart = 0;
            // *pos: in, the sorted position of the point's previous match (-1: none) — the seed of
            // the search (icp_target_nn.cuh); out, the position of this iteration's match
            auto search_one = [&](int i, double qx, double qy, double* ox, double* oy, int* pos) {
                const double pz = (DIM == 3) ? src[(size_t)i * DIM + 2] : 0.0;  // transform_xy keeps z (lib.rs:52-57)
                int idx;
                if (SHARED_TARGET) {
                    const NnResult<DIM, uint32_t, uint32_t> r = target_nn_search<DIM>(p.target, qx, qy, pz, *pos);
                    *ox = p.target.x[r.pos]; *oy = p.target.y[r.pos];
                    idx = r.idx;
                    *pos = r.pos;
                } else {
                    GridView<DIM, uint16_t, uint16_t> gv;
                    gv.x = reinterpret_cast<const double*>(U + L.nn_x);
                    gv.y = reinterpret_cast<const double*>(U + L.nn_y);
                    gv.z = reinterpret_cast<const double*>(U + L.nn_z);
                    gv.idx = reinterpret_cast<const uint16_t*>(U + L.nn_idx);
                    gv.cstart = reinterpret_cast<const uint16_t*>(U + L.nn_cell);
                    gv.g = geom;
                    NnResult<DIM, uint16_t, uint16_t> r = grid_nn_search_auto(gv, qx, qy, pz, *pos);
                    *ox = gv.x[r.pos]; *oy = gv.y[r.pos];
                    idx = r.idx;
                    *pos = r.pos;
                }
                hpart += mix64(((uint64_t)(uint32_t)i << 32) | (uint64_t)(uint32_t)idx);
                if (p.out_corr) {
                    if (p.corr_mode == 2) p.out_corr[corr_base + (uint64_t)it * (uint64_t)n + (uint64_t)i] = idx;
                    // (with the fixed-point exit any iteration can be the last one computed: write each)
                    else if (p.corr_mode == 1 && (it + 1 == p.max_iter || p.fixed_point)) p.out_corr[corr_base + (uint64_t)i] = idx;
                }
            };
            if (ORDERED) {
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const int i = tid + k * THREADS;
                    sx[k] = 0.; sy[k] = 0.; dx[k] = 0.; dy[k] = 0.;
                    if (i < n) {
                        transform_apply(T, src[(size_t)i * DIM + 0], src[(size_t)i * DIM + 1], &sx[k], &sy[k]);
                        search_one(i, sx[k], sy[k], &dx[k], &dy[k], &ppos[k]);
                    }
                }
            } else {
                // one (not PPT) inlined copy of the search: queries go through a shared scratch
                // slot per point (the thread's own), the matched target point comes back the same way
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const int i = tid + k * THREADS;
                    sx[k] = 0.; sy[k] = 0.; dx[k] = 0.; dy[k] = 0.;
                    if (i < n) {
                        transform_apply(T, src[(size_t)i * DIM + 0], src[(size_t)i * DIM + 1], &sx[k], &sy[k]);
                        double2 q;
                        q.x = sx[k]; q.y = sy[k];
                        r2v[i] = q;
                    }
                }
#pragma unroll 1
                for (int i = tid; i < n; i += THREADS) {
                    const double2 q = r2v[i];
                    double2 d;
                    int none = -1;
                    search_one(i, q.x, q.y, &d.x, &d.y, &none);
                    r2v[i] = d;
                }
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const int i = tid + k * THREADS;
                    if (i < n) {
                        const double2 d = r2v[i];
                        dx[k] = d.x; dy[k] = d.y;
                    }
                }
            }
            hpart = warp_sum_u64(hpart);
            if (lane == 0) ctl->hash_part[warp] = hpart;
            __syncthreads();  // search done (ordered: the union region becomes the IRLS row arrays)
            uint64_t hterm = 0;
            if (tid == 0) {
                for (int w = 0; w < NW; ++w) hterm += ctl->hash_part[w];
                H = H * 0x100000001B3ull + hterm;
            }

// @region irls_setup
            // ---- estimate_transform (lib.rs:59-84)
            Transform Ti = transform_identity();
            double prev_error = kF64Max;
            uint32_t accepted = 0, why = kBreakCap;
            if (n < 2) {
                why = kBreakNone;  // check_input_size (lib.rs:186-189,225-228)
            } else {
                for (int inner = 0; inner < kInnerMaxIter; ++inner) {
                    double sigma[2];
                    double jtj[9], jtr[3], error;
                    if (ORDERED) {
// @region ordered_rows
                        // residuals (lib.rs:230-234)
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            const int i = tid + k * THREADS;
                            if (i < npad) {
                                double2 rr;
                                rr.x = 0.; rr.y = 0.;
                                if (i < n) {
                                    double tx_, ty_;
                                    transform_apply(Ti, sx[k], sy[k], &tx_, &ty_);
                                    rr.x = tx_ - dx[k];
                                    rr.y = ty_ - dy[k];
                                }
                                r2v[i] = rr;
                            }
                        }
                        if (tid < 2 * kChainUnroll) {
                            double2 kv;
                            if (tid < kChainUnroll) { kv.x = Ti.r00; kv.y = Ti.r10; }
                            else { kv.x = Ti.r01; kv.y = Ti.r11; }
                            ktab[tid < kChainUnroll ? tid : tid - kChainUnroll + kChainK1] = kv;
                        }
                        __syncthreads();
                        // calc_stddevs (stats.rs:49-60): per-axis median, then median of |r - median|
                        KeysInShared ks;
                        ks.v = r2v;
                        hist_par = cta_calc_stddevs<THREADS, PPT>(sel, ks, n, est_med, est_sigma, hist_par, sigma);
                        // row weights (lib.rs:245-250), Jacobian third column (lib.rs:176-184),
                        // Huber cost terms (lib.rs:47-48)
                        const double g0 = 1. / sigma[0], g1 = 1. / sigma[1];
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            const int i = tid + k * THREADS;
                            if (i < npad) {
                                double2 cc, bb;
                                double rh = 0.;
                                cc.x = 0.; cc.y = 0.; bb.x = 0.; bb.y = 0.;
                                if (i < n) {
                                    const double2 rr = r2v[i];
                                    if (sigma[0] != 0.) cc.x = huber_drho_of_residual(rr.x) * g0;
                                    if (sigma[1] != 0.) cc.y = huber_drho_of_residual(rr.y) * g1;
                                    const double a0 = -sy[k], a1 = sx[k];
                                    bb.x = Ti.r00 * a0 + Ti.r01 * a1;
                                    bb.y = Ti.r10 * a0 + Ti.r11 * a1;
                                    rh = huber_rho(rr.x * rr.x + rr.y * rr.y);
                                }
                                c2v[i] = cc;
                                b2v[i] = bb;
                                rho[i] = rh;
                            }
                        }
                        __syncthreads();
                        // ordered accumulation
                        if (warp == chain_warp) {
                            const double s = chain_run(chain_role(lane, r2v, c2v, b2v, ktab), npad / kChainStage);
                            if (lane < 12) ctl->sums[lane] = s;
                        }
                        if (warp == cost_warp) {
                            const double e = cost_run(smem_ref_of(reinterpret_cast<const double2*>(rho)), npad / 2);
                            if (lane == 0) ctl->sums[12] = e;
                        }
                        __syncthreads();
#pragma unroll
                        for (int q = 0; q < 9; ++q) jtj[q] = ctl->sums[q];
#pragma unroll
                        for (int q = 0; q < 3; ++q) jtr[q] = ctl->sums[9 + q];
                        error = ctl->sums[12];
                    } else {
// @region tree_residuals
                        // residuals (lib.rs:230-234) stay in registers
                        double2 rr[PPT];
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            double tx_, ty_;
                            transform_apply(Ti, sx[k], sy[k], &tx_, &ty_);
                            rr[k].x = tx_ - dx[k];
                            rr[k].y = ty_ - dy[k];
                        }
                        KeysInRegisters<PPT> ks;
                        ks.v = rr;
                        hist_par = cta_calc_stddevs<THREADS, PPT>(sel, ks, n, est_med, est_sigma, hist_par, sigma);
// @region tree_accumulate
                        // row weights (lib.rs:245-250), Jacobian (lib.rs:176-184), Huber cost (lib.rs:47-48)
                        // and their sums, reduced as a fixed tree: per-thread partials over the
                        // thread's points (ascending), split butterfly across lanes, warps in order
                        const double g0 = 1. / sigma[0], g1 = 1. / sigma[1];
                        double acc[16];   // 0-5 upper triangle of jtj, 6-8 jtr, 9 cost, 10-15 unused (zero)
#pragma unroll
                        for (int q = 0; q < 16; ++q) acc[q] = 0.;
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            const int i = tid + k * THREADS;
                            if (i < n) {
                                tree_accumulate_point(acc, Ti, sx[k], sy[k], rr[k], sigma, g0, g1);
                            }
                        }
// @region tree_reduce
                        cta_tree_reduce_solve<NW>(ctl, acc, Ti, prev_error, p.guard);
                        const uint32_t decision = (uint32_t)ctl->bcast[7];
                        if (decision != kBreakCap) { why = decision; break; }
                        prev_error = ctl->bcast[6];
                        Ti.r00 = ctl->bcast[0]; Ti.r01 = ctl->bcast[1]; Ti.r10 = ctl->bcast[2];
                        Ti.r11 = ctl->bcast[3]; Ti.tx = ctl->bcast[4];  Ti.ty = ctl->bcast[5];
                        ++accepted;
                        continue;
                    }
// @region solve_update
                    // solve + break tests + update: every thread, identically (lib.rs:67-81)
                    double delta[3];
                    if (!solve_normal_equations(jtj, jtr, delta)) { why = kBreakNone; break; }
                    if ((delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2] < kDeltaNormThreshold) {
                        why = kBreakDelta;
                        break;
                    }
                    if (error > prev_error) { why = kBreakError; break; }
                    prev_error = error;
                    Ti = transform_mul(transform_new(delta[0], delta[1], delta[2]), Ti);
                    ++accepted;
                }
            }
// @region outer_tail
            const bool repeats = fused_repeats(p, accepted, it);
            if (p.out_cost && (it + 1 == p.max_iter || repeats)) {
                // huber_error (lib.rs:45-50) of the inner solve's final transform over this (the last)
                // iteration's transformed source / matched target points
                __syncthreads();   // every thread has read the last solve's sums / broadcast
                if (ORDERED) {
#pragma unroll
                    for (int k = 0; k < PPT; ++k) {
                        const int i = tid + k * THREADS;
                        if (i < npad) {
                            double rh = 0.;
                            if (i < n) {
                                double tx_, ty_;
                                transform_apply(Ti, sx[k], sy[k], &tx_, &ty_);
                                const double rx = tx_ - dx[k], ry = ty_ - dy[k];
                                rh = huber_rho(rx * rx + ry * ry);
                            }
                            rho[i] = rh;
                        }
                    }
                    __syncthreads();
                    if (warp == cost_warp) {
                        const double e = npad > 0 ? cost_run(smem_ref_of(reinterpret_cast<const double2*>(rho)), npad / 2) : 0.;
                        if (lane == 0) p.out_cost[pair] = e;
                    }
                } else {
                    double e = 0.;
#pragma unroll
                    for (int k = 0; k < PPT; ++k) {
                        const int i = tid + k * THREADS;
                        double tx_, ty_;
                        transform_apply(Ti, sx[k], sy[k], &tx_, &ty_);
                        const double rx = tx_ - dx[k], ry = ty_ - dy[k];
                        const double rh = huber_rho(rx * rx + ry * ry);
                        e = e + (i < n ? rh : 0.);
                    }
                    const double tot = cta_tree_sum<NW>(ctl, e);
                    if (tid == 0) p.out_cost[pair] = tot;
                }
            }
            T = transform_mul(Ti, T);  // lib.rs:127
            if (tid == 0 && p.out_inner) p.out_inner[(size_t)pair * p.max_iter + it] = accepted | (why << 16);
            if (repeats) {
                fused_repeat_outputs<THREADS>(p, pair, it, n, corr_base, accepted | (why << 16));
                if (tid == 0) H = fused_repeat_hash(H, hterm, p.max_iter - 1u - it);
            }
            __syncthreads();  // rows are dead; next iteration restages the target
            if (repeats) break;
        }
        if (tid == 0) {
            p.out_T[(size_t)pair * 6 + 0] = T.r00; p.out_T[(size_t)pair * 6 + 1] = T.r01;
            p.out_T[(size_t)pair * 6 + 2] = T.r10; p.out_T[(size_t)pair * 6 + 3] = T.r11;
            p.out_T[(size_t)pair * 6 + 4] = T.tx;  p.out_T[(size_t)pair * 6 + 5] = T.ty;
            if (p.out_status) p.out_status[pair] = kPairOk;
            if (p.out_hash) p.out_hash[pair] = H;
            if (p.out_cost && p.max_iter == 0) p.out_cost[pair] = 0.;
            if (!ORDERED) fused_report_knife(p, ctl, pair);
        }
        __syncthreads();  // ctl->pair is rewritten at the top of the loop
    }
    if (tid == 0) fused_launch_exit(p);
}

}  // namespace icp
