// icp_target_nn.cuh — seeded correspondence search: on a large shared target (the `Icp2d<'a>` value
// kept on the device, src/lib.rs:91-102; KdTree::search per query, src/lib.rs:121) and, seeded with
// the previous outer iteration's match, on per-pair grids (grid_nn_search_seeded, end of file).
//
// Contract (identical to icp_grid.cuh and the oracle's brute force): exact Euclidean 1-NN on the
// COMPUTED d2 = dx*dx + dy*dy (binary64, dx = q.x - p.x, no FMA), lowest original index on equal d2.
//
// 2-D targets carry, besides the sorted cell list of icp_grid.cuh,
//   seed[c]  u32     sorted position of a target point near cell c (jump flooding at build time);
//                    ANY point is a valid seed — its exact d2 bounds the search disc, nothing more
//   pf[p]    float2  binary32 copy of point p in CELL UNITS relative to the centre of its 32x32-cell
//                    block: (T(p.x) - centre, T(p.y) - centre) with T(v) = (v - o) * inv_h, the very
//                    expression grid_coord floors.  |.| <= 17, so one rounding is <= 2^-24 * 17 cells
// and the search runs
//   1. U2 = exact d2 of the seed: every point the disc test skips has computed d2 > U2;
//   2. one row of cells at a time, the cells the disc of radius sqrt(U2) reaches (a contiguous span
//      of the sorted arrays), binary32 distances only: smallest, its position, second smallest;
//   3. the smallest is the answer outright when the second smallest lies outside the proven error
//      band of the filter; otherwise (ties, near-ties, duplicates) the spans are walked again and
//      every candidate inside the band is ranked on its exact binary64 d2 / original index.
// Every skip is justified by a bound that holds for the computed binary64 d2 (comments "bound:"),
// so the result is bit-identical to scanning all target points.  Geometry, seeds and the filter
// change speed only.  3-D targets and degenerate geometries use the ring search of icp_grid.cuh.
//
// One thread per query, in nn_search_kernel and in the fused kernels alike.  A warp-collective form
// that pooled the candidates of 32 queries in shared memory (chunks of 8, folded into their owners'
// slots with shared-memory atomics) ran the candidate loop at 30 of 32 lanes but lost more in list
// upkeep than it won: 4.7 against 5.5 G queries/s on B200 (profiles/r01_nn_pooled_variant.jsonl,
// profiles/r01_ncu_full_summary.md); removed.
#pragma once
#include "icp_grid.cuh"

namespace icp {

// Shared target (Icp2d::new analogue kept on the device, reused by many estimates)
struct TargetView {
    const double* x;
    const double* y;
    const double* z;
    const uint32_t* idx;
    const uint32_t* cstart;
    const uint32_t* seed;   // 2-D fast path only, else null
    const float2* pf;       // 2-D fast path only, else null
    GridGeom g;
    uint32_t m;
};

constexpr int kNnBlockShift = 5;                    // 32x32-cell blocks share one binary32 origin
constexpr int kNnBlockCells = 1 << kNnBlockShift;
constexpr uint32_t kNnNoSeed = 0xffffffffu;

// Centre of the block that holds cell c, in cell units.  Build and search evaluate this same expression.
ICP_HD double nn_block_centre(int c)
{
    return (double)(((c >> kNnBlockShift) << kNnBlockShift) + kNnBlockCells / 2);
}

// T(v): the cell coordinate grid_coord floors (before its clamp).  Monotone non-decreasing in v.
ICP_HD double nn_cell_coord(double v, double o, double inv_h) { return (v - o) * inv_h; }

ICP_DEV float nn_fma(float a, float b, float c)
{
#ifdef ICP_SIMT_EMULATE
    return std::fmaf(a, b, c);
#else
    return __fmaf_rn(a, b, c);
#endif
}

// floor to int, clamped to [0, hi]; |v| <= ~1e9 is guaranteed by the caller (NnQuery::slow otherwise)
ICP_DEV int nn_floor_clamp(float v, int hi)
{
#ifdef ICP_SIMT_EMULATE
    const int c = (int)std::floor(v);
#else
    const int c = __float2int_rd(v);
#endif
    return c < 0 ? 0 : (c > hi ? hi : c);
}

#ifdef ICP_SIMT_EMULATE
// emulator-only work counters (tests/simt): queries, rows, spans, candidates of pass 1, pass-2 queries
struct NnCounters { unsigned long long q, rows, spans, cand, pass2, exact; };
inline NnCounters& nn_counters() { static NnCounters c{}; return c; }
#define ICP_NN_COUNT(field, v) (nn_counters().field += (v))
#else
#define ICP_NN_COUNT(field, v) ((void)0)
#endif

struct NnFilterState {
    float m1, m2;   // smallest and second smallest binary32 d2 seen (cell units squared)
    int pos;        // sorted position of the smallest
};

// Per-query constants of the disc walk.  All binary32 quantities are in cell units.
//
// bound (cells): a point p of column c has T(p.x) in [c, c+1) up to the clamp at the grid's edge
// (where no points lie beyond), T is the monotone map above, and T(v) = (v - o)/h (1 + e), |e| <
// 2^-51.  So for a column c > x: (p.x - q.x)/h >= c - T(q.x) - 2^-50 gx, and likewise on the other
// side and for rows.  txf/tyf are T(q) rounded to binary32; `sx`/`sy` = 1.5 * 2^-22 (|t| + 1) cover that
// rounding, the rounding of every binary32 difference formed with them, and the 2^-50 terms.
struct NnQuery {
    double tx, ty;     // T(q.x), T(q.y)
    float txf, tyf, sx, sy;
    float Uc2;         // >= U2 (1 + 1e-6) / h^2
    float a;           // error bound of one binary32 coordinate difference
    int ya, yb;        // rows the disc reaches
    uint32_t seed;
    double U2;         // exact d2 of the seed
    bool slow;         // outside the filter's range: ring search instead
};

// seed_pos < 0: the seed of the query's cell; else any sorted position (e.g. the match the same
// source point had in the previous outer iteration — usually the tighter disc)
ICP_DEV NnQuery nn_query_setup(const TargetView& tv, double qx, double qy, int seed_pos = -1)
{
    const GridGeom& g = tv.g;
    NnQuery Q;
    Q.tx = nn_cell_coord(qx, g.ox, g.inv_h);
    Q.ty = nn_cell_coord(qy, g.oy, g.inv_h);
    Q.txf = (float)Q.tx; Q.tyf = (float)Q.ty;
    Q.slow = false;
    Q.ya = 1; Q.yb = 0; Q.Uc2 = 0.0f; Q.a = 0.0f; Q.sx = 0.0f; Q.sy = 0.0f; Q.U2 = 0.0;
    if (seed_pos >= 0) {
        Q.seed = (uint32_t)seed_pos;
    } else {
        const int cx = grid_coord(qx, g.ox, g.inv_h, g.gx);
        const int cy = grid_coord(qy, g.oy, g.inv_h, g.gy);
        Q.seed = tv.seed[cy * g.gx + cx];
    }
    if (Q.seed >= tv.m || !(fabsf(Q.txf) + fabsf(Q.tyf) <= 1e9f)) { Q.slow = true; return Q; }
    const double dx = qx - tv.x[Q.seed], dy = qy - tv.y[Q.seed];
    Q.U2 = dx * dx + dy * dy;
    // binary32 upper bound of U2 / h^2: four roundings of 2^-24 against the factor 1.000001, and
    // 1e-30 where binary32 would go subnormal
    const float ihf = (float)g.inv_h;
    Q.Uc2 = ((float)Q.U2 * ihf) * ihf * 1.000001f + 1e-30f;
    if (!(Q.Uc2 <= 1e12f) || !(Q.U2 >= 0.0)) { Q.slow = true; return Q; }
    const float Uc = sqrtf(Q.Uc2) * 1.000001f;
    Q.sx = (fabsf(Q.txf) + 1.0f) * 3.6e-7f;   // 1.5 * 2^-22
    Q.sy = (fabsf(Q.tyf) + 1.0f) * 3.6e-7f;
    // bound (rows): a point in a row above yb has (p.y - q.y)/h > Uc, hence |q.y - p.y| >
    // sqrt(U2)(1 + 4e-7) and its computed d2 exceeds U2.  Same below ya.
    Q.ya = nn_floor_clamp(Q.tyf - Uc - Q.sy, g.gy - 1);
    Q.yb = nn_floor_clamp(Q.tyf + Uc + Q.sy, g.gy - 1);
    // the stored point (|.| <= 17), the query's block offset (|.| <= Uc + 18) and their difference
    // (|.| <= Uc + 35) are each rounded to 24 bits
    Q.a = (Uc + 40.0f) * 2.4e-7f;
    return Q;
}

// Columns of row yy the disc reaches: false when none.
// bound (columns): points of row yy have |q.y - p.y|/h >= ydc; outside [T(q.x) - wc, T(q.x) + wc]
// their computed d2 is >= h^2 (wc^2 + ydc^2)(1 - 2^-49) >= U2 (1 + 5e-7) > U2.
ICP_DEV bool nn_row_span(const NnQuery& Q, int yy, int gx, int* xa, int* xb)
{
    const float yf = (float)yy;
    const float ydc = fmaxf(fmaxf(yf - Q.tyf, Q.tyf - (yf + 1.0f)) - Q.sy, 0.0f) * 0.999999f;
    const float w2 = Q.Uc2 - ydc * ydc;
    if (w2 < 0.0f) return false;
    const float wc = sqrtf(w2) * 1.000001f + Q.sx;
    *xa = nn_floor_clamp(Q.txf - wc, gx - 1);
    *xb = nn_floor_clamp(Q.txf + wc, gx - 1);
    return true;
}

// query offsets from the centre of the block holding column / row c, rounded once to binary32
ICP_DEV float nn_block_offset(double t, int c) { return (float)(t - nn_block_centre(c)); }

// bound (filter): sqrt(d32) is within 1.5a + 2^-22 sqrt(d32) of the true distance D/h.  The winner w
// of the exact ranking has D(w) <= D(p)(1 + 2^-50) for every p, in particular for the binary32
// minimum, hence sqrt(d32(w)) <= sqrt(m1)(1 + 2^-20) + 4a.  Candidates above the band cannot win.
ICP_DEV float nn_filter_band(float m1, float a)
{
    const float t = sqrtf(m1) * 1.000004f + 5.0f * a;
    return t * t * 1.000001f + 1e-30f;
}

ICP_DEV NnResult<2, uint32_t, uint32_t> nn_ring_search(const TargetView& tv, double qx, double qy)
{
    GridView<2, uint32_t, uint32_t> gv;
    gv.x = tv.x; gv.y = tv.y; gv.z = tv.z; gv.idx = tv.idx; gv.cstart = tv.cstart; gv.g = tv.g;
    return grid_nn_search(gv, qx, qy, 0.0);
}

ICP_DEV NnResult<2, uint32_t, uint32_t> nn_result_at(const TargetView& tv, double qx, double qy, int p)
{
    NnResult<2, uint32_t, uint32_t> r;
    const double dx = qx - tv.x[p], dy = qy - tv.y[p];
    r.d2 = dx * dx + dy * dy;
    r.pos = p;
    r.idx = (int)tv.idx[p];
    return r;
}

// Walks the cells the disc reaches; `visit(p, d32)` sees every candidate's binary32 d2.
// Shape of the loop: each trip first advances to the query's next NON-EMPTY span (cells of one row
// inside one 32-cell block), then runs the candidates of that span.  The lanes of a warp meet their
// occupied rows at different row numbers; advancing first lets them run their k-th candidate loops
// together instead of one lane at a time.
// SHRINK: a candidate closer than the seed shrinks the disc for the rows still ahead.  The closest
// candidate so far has a true distance of at most (sqrt(m1)(1 + 2^-22) + 1.5a) h (bound (filter));
// a point outside the disc of that radius is farther than it by more than any rounding involved, so
// it is neither the winner nor able to tie with it, and the filter's "single candidate inside the
// band" test stays a test over every point that could win.  (The exact pass walks the seed's disc.)
template <bool SHRINK = false, typename Visit>
ICP_DEV void nn_walk(const TargetView& tv, const NnQuery& Qin, Visit&& visit)
{
    NnQuery Q = Qin;
    float m1w = 3.0e38f;
    const int gx = tv.g.gx;
    int yy = Q.ya, xs = 1, xb = 0, row = 0;
    float qry = 0.0f;
    bool have_row = false;
    for (;;) {
        int p0 = 0, p1 = 0;
        float qrx = 0.0f;
        while (p1 == p0 && (have_row || yy <= Q.yb)) {
            if (!have_row) {
                int xa;
                if (!nn_row_span(Q, yy, gx, &xa, &xb)) { ++yy; continue; }
                ICP_NN_COUNT(rows, 1);
                row = yy * gx;
                const int r0 = (int)tv.cstart[row + xa], r1 = (int)tv.cstart[row + xb + 1];
                if (r0 == r1) { ++yy; continue; }   // nothing in the whole span
                qry = nn_block_offset(Q.ty, yy);
                xs = xa;
                if ((xa >> kNnBlockShift) == (xb >> kNnBlockShift)) {   // one block: the span is the segment
                    p0 = r0; p1 = r1;
                    qrx = nn_block_offset(Q.tx, xa);
                    ICP_NN_COUNT(spans, 1);
                    ++yy;
                    break;
                }
                have_row = true;
            }
            const int xe_blk = xs | (kNnBlockCells - 1);
            const int xe = xe_blk < xb ? xe_blk : xb;
            p0 = (int)tv.cstart[row + xs];
            p1 = (int)tv.cstart[row + xe + 1];
            qrx = nn_block_offset(Q.tx, xs);
            ICP_NN_COUNT(spans, 1);
            xs = xe + 1;
            if (xs > xb) { have_row = false; ++yy; }
        }
        if (p1 == p0) break;
        ICP_NN_COUNT(cand, p1 - p0);
        // candidates in ascending position, two per 16-byte load where the position is even (pf is
        // 16-byte aligned): a lane's fetches are the L1/TEX unit's work — one wavefront per load.
        // (Four per 256-bit load — sm_100's LDG.E.256 — was measured: 4.79 against 6.33 G queries/s.  Two
        // 16-byte loads in flight per trip: 5.03 G at the 80 registers it takes, 5.79 G held to 64; config 3
        // 49.9k against 55.3k scans/s.  The walk is bound by L1/TEX wavefronts and issue slots, not by the
        // latency of its one load in flight.)
        int p = p0;
        const float m1_before = m1w;
        if ((p & 1) != 0) {
            const float2 v = tv.pf[p];
            const float ex = qrx - v.x, ey = qry - v.y;
            const float d = nn_fma(ex, ex, ey * ey);
            if (SHRINK) m1w = fminf(m1w, d);
            visit(p, d);
            ++p;
        }
        for (; p + 1 < p1; p += 2) {
            const float4 w = *reinterpret_cast<const float4*>(tv.pf + p);
            const float ex0 = qrx - w.x, ey0 = qry - w.y, ex1 = qrx - w.z, ey1 = qry - w.w;
            const float d0 = nn_fma(ex0, ex0, ey0 * ey0), d1 = nn_fma(ex1, ex1, ey1 * ey1);
            if (SHRINK) m1w = fminf(m1w, fminf(d0, d1));
            visit(p, d0);
            visit(p + 1, d1);
        }
        if (p < p1) {
            const float2 v = tv.pf[p];
            const float ex = qrx - v.x, ey = qry - v.y;
            const float d = nn_fma(ex, ex, ey * ey);
            if (SHRINK) m1w = fminf(m1w, d);
            visit(p, d);
        }
        if (SHRINK && m1w < m1_before) {
            const float uc = sqrtf(m1w) * 1.000001f + 2.0f * Q.a;
            const float u2 = uc * uc * 1.000002f;
            if (u2 < Q.Uc2) {
                Q.Uc2 = u2;
                const int yb2 = nn_floor_clamp(Q.tyf + uc * 1.000002f + Q.sy, tv.g.gy - 1);   // bound (rows)
                Q.yb = yb2 < Q.yb ? yb2 : Q.yb;
            }
        }
    }
}

// Exact ranking of the candidates inside the band (ties, near-ties, duplicates).
ICP_DEV NnResult<2, uint32_t, uint32_t> nn_exact_pass(const TargetView& tv, const NnQuery& Q, double qx, double qy,
                                                      float thr)
{
    NnResult<2, uint32_t, uint32_t> best;
    best.d2 = Q.U2; best.pos = (int)Q.seed; best.idx = (int)tv.idx[Q.seed];
    ICP_NN_COUNT(pass2, 1);
    nn_walk(tv, Q, [&](int p, float d32) {
        if (d32 <= thr) {
            ICP_NN_COUNT(exact, 1);
            const double dx = qx - tv.x[p], dy = qy - tv.y[p];
            const double d2 = dx * dx + dy * dy;
            if (d2 <= best.d2) {
                const int id = (int)tv.idx[p];
                if (d2 < best.d2 || id < best.idx) { best.d2 = d2; best.idx = id; best.pos = p; }
            }
        }
    });
    return best;
}

// Exact 1-NN on a 2-D shared target with seeds.  Requires tv.m > 0, tv.seed, tv.pf.
// need_d2 false (the stand-alone search kernel wants the index only): the winner's exact d2 is not formed,
// which saves its two 8-byte coordinate fetches — the kernel is bound by L1/TEX wavefronts, one per fetch.
ICP_DEV NnResult<2, uint32_t, uint32_t> target_nn_search2(const TargetView& tv, double qx, double qy, int seed_pos = -1,
                                                          bool need_d2 = true)
{
    const NnQuery Q = nn_query_setup(tv, qx, qy, seed_pos);
    if (Q.slow) return nn_ring_search(tv, qx, qy);
    ICP_NN_COUNT(q, 1);
    NnFilterState f;
    f.m1 = 3.0e38f; f.m2 = 3.0e38f; f.pos = (int)Q.seed;
    // (nn_walk<true>, the disc shrinking as closer candidates turn up, was measured on the search
    // kernel: 5.34 against 6.28 G queries/s, config 3 56.1k against 58.0k scans/s — the walk reaches
    // the near candidates halfway through its rows, and the bookkeeping costs more than the rest saves)
    nn_walk(tv, Q, [&](int p, float d32) {
        f.m2 = fminf(f.m2, fmaxf(d32, f.m1));
        f.pos = (d32 < f.m1) ? p : f.pos;
        f.m1 = fminf(f.m1, d32);
    });
    const float thr = nn_filter_band(f.m1, Q.a);
    // a single candidate inside the band, and the exact winner is inside the band: it is the winner
    if (f.m2 > thr && f.m1 < 3.0e38f) {
        if (need_d2) return nn_result_at(tv, qx, qy, f.pos);
        NnResult<2, uint32_t, uint32_t> r;
        r.d2 = 0.0; r.pos = f.pos; r.idx = (int)tv.idx[f.pos];
        return r;
    }
    return nn_exact_pass(tv, Q, qx, qy, thr);
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative form of the same search, for COHERENT queries: the 32 lanes of a warp hold 32
// queries that lie close together (consecutive beams of a scan hit the same walls a few
// millimetres apart — the queries ICP issues).  The discs of the 32 queries overlap almost
// entirely, so instead of 32 private walks (each lane fetching its own 8-byte candidates and 4-byte
// cell-table entries: the L1/TEX unit saturates at ~40 sectors per query and the lanes idle on the
// heavy tail of the candidate count) the warp walks the UNION of the discs once:
//   rows      every lane forms its own chord of the row (nn_row_span, same bounds as above); two
//             warp reductions give the union chord; its two cell-table entries are one broadcast
//             load for the whole warp;
//   candidates of the union chord are fetched 32 at a time with ONE coalesced 256-byte load,
//             staged in shared memory (double buffered, one __syncwarp per chunk) and read back as
//             broadcasts: every lane forms its binary32 d2 to every staged candidate with all 32
//             lanes active, in the same arithmetic as the private walk (offsets from the centre of
//             the candidate's 32x32-cell block).
// A lane therefore sees a SUPERSET of its own candidates.  Nothing in the filter's proof needs the
// candidates to lie inside the lane's disc — only the error bound of one binary32 coordinate
// difference, which grows with the largest difference formed: `a` is taken for the whole frame
// (W cells) instead of the lane's own radius.  The frame is limited to kNnWarpExtent cells; a warp
// whose queries are spread wider (incoherent queries), or that holds a query outside the filter's
// range, falls back to the private search for all its lanes.  Ties, near-ties and duplicates go to
// the lane's own exact pass as before.  Results are bit-identical to target_nn_search2.
constexpr int kNnWarpExtent = 96;

struct NnWarpScratch {
    float2 c[2][32];   // one chunk of candidates, double buffered
};

ICP_DEV int nn_warp_min(int v) { return __reduce_min_sync(ICP_FULL_MASK, v); }
ICP_DEV int nn_warp_max(int v) { return __reduce_max_sync(ICP_FULL_MASK, v); }

// Must be called by all 32 lanes of a converged warp (padding lanes repeat a real query).
// Returns the sorted position of the exact nearest target point of (qx, qy).
ICP_DEV int target_nn_search2_warp(const TargetView& tv, double qx, double qy, NnWarpScratch* ws)
{
    const NnQuery Q = nn_query_setup(tv, qx, qy);
    const int lane = lane_id();
    const int gx = tv.g.gx;
    int xlo = 0, xhi = 0;
    if (!Q.slow) {
        const float Uc = sqrtf(Q.Uc2) * 1.000001f;
        xlo = nn_floor_clamp(Q.txf - Uc - Q.sx, gx - 1);
        xhi = nn_floor_clamp(Q.txf + Uc + Q.sx, gx - 1);
    }
    const bool any_slow = __any_sync(ICP_FULL_MASK, Q.slow);
    const int Y0 = nn_warp_min(Q.ya), Y1 = nn_warp_max(Q.yb), X0 = nn_warp_min(xlo), X1 = nn_warp_max(xhi);
    const int W = ((X1 - X0) > (Y1 - Y0) ? (X1 - X0) : (Y1 - Y0)) + 1;
    if (any_slow || W > kNnWarpExtent) return target_nn_search2(tv, qx, qy).pos;   // incoherent warp: private walks
    ICP_NN_COUNT(q, 1);
    // the stored point (|.| <= 17), the query's block offset (|.| <= W + 17) and their difference
    // (|.| <= W + 35) are each rounded to 24 bits: see NnQuery::a
    const float a = ((float)W + 40.0f) * 2.4e-7f;
    NnFilterState f;
    f.m1 = 3.0e38f; f.m2 = 3.0e38f; f.pos = (int)Q.seed;
    int buf = 0;
    for (int yy = Y0; yy <= Y1; ++yy) {
        int xa = 0x7fffffff, xb = -1;
        if (yy >= Q.ya && yy <= Q.yb) {
            int ca, cb;
            if (nn_row_span(Q, yy, gx, &ca, &cb)) { xa = ca; xb = cb; }   // bound (columns), per lane
        }
        const int Xa = nn_warp_min(xa), Xb = nn_warp_max(xb);
        if (Xa > Xb) continue;
        ICP_NN_COUNT(rows, 1);
        const int row = yy * gx;
        const float qry = nn_block_offset(Q.ty, yy);
        for (int xs = Xa; xs <= Xb;) {
            const int xe_blk = xs | (kNnBlockCells - 1);
            const int xe = xe_blk < Xb ? xe_blk : Xb;
            const int p0 = (int)tv.cstart[row + xs], p1 = (int)tv.cstart[row + xe + 1];   // the same address in every lane
            if (p1 > p0) {
                ICP_NN_COUNT(spans, 1);
                ICP_NN_COUNT(cand, p1 - p0);
                const float qrx = nn_block_offset(Q.tx, xs);
                for (int base = p0; base < p1; base += 32) {
                    const int cnt = (p1 - base) < 32 ? (p1 - base) : 32;
                    float2 v;
                    v.x = 0.0f; v.y = 0.0f;
                    if (lane < cnt) v = tv.pf[base + lane];     // one coalesced load per chunk
                    ws->c[buf][lane] = v;
                    __syncwarp();
#pragma unroll 4
                    for (int j = 0; j < cnt; ++j) {
                        const float2 c = ws->c[buf][j];         // broadcast
                        const float ex = qrx - c.x, ey = qry - c.y;
                        const float d32 = nn_fma(ex, ex, ey * ey);
                        f.m2 = fminf(f.m2, fmaxf(d32, f.m1));
                        f.pos = (d32 < f.m1) ? base + j : f.pos;
                        f.m1 = fminf(f.m1, d32);
                    }
                    buf ^= 1;
                }
            }
            xs = xe + 1;
        }
    }
    const float thr = nn_filter_band(f.m1, a);
    if (f.m2 > thr && f.m1 < 3.0e38f) return f.pos;
    return nn_exact_pass(tv, Q, qx, qy, thr).pos;
}

// ---------------------------------------------------------------------------------------------
// The same search by a GROUP of G lanes per query (G = 4 or 8; 32 / G queries per warp at a time).
// The private walk is bound by the L1/TEX unit: a lane's every fetch is a wavefront of its own
// (~47 per query: ~12 cell-table entries, ~30 candidate pairs, the seed) and only ~13 of 32 lanes
// are active in the candidate loop.  Here the lanes of a group share ONE query:
//   rows        lane j of the group takes rows ya + j, ya + j + G, ... of the disc: its chord
//               (nn_row_span), the chord's two cell-table entries, one segment per 32-cell block;
//   candidates  the segments are handed round the group (two shuffles each) and scanned by all G
//               lanes, two candidates per lane and step from consecutive 16-byte words — G words
//               are one wavefront — each lane keeping its own filter state (m1, m2, position);
//   finish      shuffle-merge of the filter states (the second smallest of a union is
//               min(m2, m2', max(m1, m1'))), band test, and for a query with two or more
//               candidates inside the band the exact pass, run by the group's first lane.
// Same bounds, same filter, same exact pass as target_nn_search2: identical results.  The loops are
// warp-uniform (a group that has run out of rows or candidates idles until the slowest is done):
// queries of one warp are consecutive entries of the query array — consecutive beams of a scan.
// Call with all 32 lanes; the lanes of a group pass the same query; `active` false for a group
// without one.  Returns the sorted position of the nearest point (group-uniform), -1 if !active.
// MEASURED (B200, 16.8 M scan queries on the 100k-point map): 4.08 G queries/s with G = 4, 3.31 G with
// G = 8, against 6.41 G for the private walk (random queries: 4.8 / 3.8 against 10.1 G): the hand-round
// of segments and the warp-uniform loops cost more than the shared wavefronts save.  Not the default.
template <int G>
ICP_DEV int target_nn_search2_group(const TargetView& tv, double qx, double qy, bool active, int seed_pos = -1)
{
    const int lane = lane_id();
    const int gl = lane & (G - 1), gbase = lane & ~(G - 1);
    const int gx = tv.g.gx;
    NnQuery Q;
    Q.slow = true; Q.ya = 1; Q.yb = 0; Q.seed = 0u; Q.a = 0.0f; Q.tx = 0.; Q.ty = 0.;
    Q.txf = 0.0f; Q.tyf = 0.0f; Q.sx = 0.0f; Q.sy = 0.0f; Q.Uc2 = 0.0f; Q.U2 = 0.;
    if (active) Q = nn_query_setup(tv, qx, qy, seed_pos);   // seed_pos >= 0: e.g. the previous outer iteration's match
    const bool walk = active && !Q.slow;
    NnFilterState f;
    f.m1 = 3.0e38f; f.m2 = 3.0e38f; f.pos = (int)Q.seed;
    auto visit = [&](int p, float d32) {
        f.m2 = fminf(f.m2, fmaxf(d32, f.m1));
        f.pos = (d32 < f.m1) ? p : f.pos;
        f.m1 = fminf(f.m1, d32);
    };
    const int NY = __reduce_max_sync(ICP_FULL_MASK, walk ? Q.yb - Q.ya + 1 : 0);
    for (int r0 = 0; r0 < NY; r0 += G) {
        // this lane's row of the disc
        const int yy = Q.ya + r0 + gl;
        int xa = 0, xb = -1;
        bool row = walk && yy <= Q.yb && nn_row_span(Q, yy, gx, &xa, &xb);
        const int rowbase = yy * gx;
        int rp0 = 0, rp1 = 0;
        if (row) {
            ICP_NN_COUNT(rows, 1);
            rp0 = (int)tv.cstart[rowbase + xa];
            rp1 = (int)tv.cstart[rowbase + xb + 1];
            row = rp1 > rp0;            // nothing in the whole chord
        }
        const float qry = row ? nn_block_offset(Q.ty, yy) : 0.0f;
        int xs = xa;
        while (__any_sync(ICP_FULL_MASK, row)) {
            // this lane's next segment: the part of its chord inside one 32-cell block
            int p0 = 0, p1 = 0;
            float qrx = 0.0f;
            if (row) {
                const int xe_blk = xs | (kNnBlockCells - 1);
                const int xe = xe_blk < xb ? xe_blk : xb;
                p0 = (xs == xa) ? rp0 : (int)tv.cstart[rowbase + xs];
                p1 = (xe == xb) ? rp1 : (int)tv.cstart[rowbase + xe + 1];
                qrx = nn_block_offset(Q.tx, xs);
                ICP_NN_COUNT(spans, 1);
                xs = xe + 1;
                row = xs <= xb;
            }
#pragma unroll
            for (int o = 0; o < G; ++o) {
                const int sp0 = __shfl_sync(ICP_FULL_MASK, p0, gbase + o);
                const int sp1 = __shfl_sync(ICP_FULL_MASK, p1, gbase + o);
                if (!__any_sync(ICP_FULL_MASK, sp1 > sp0)) continue;
                const float sqrx = __uint_as_float(__shfl_sync(ICP_FULL_MASK, __float_as_uint(qrx), gbase + o));
                const float sqry = __uint_as_float(__shfl_sync(ICP_FULL_MASK, __float_as_uint(qry), gbase + o));
                if (gl == 0) ICP_NN_COUNT(cand, sp1 - sp0);
                // two candidates per lane and step; word w holds positions 2w and 2w + 1 (pf is 16-byte aligned)
                for (int b = sp0 & ~1; __any_sync(ICP_FULL_MASK, b < sp1); b += 2 * G) {
                    const int p = b + 2 * gl;
                    if (p + 1 < sp1) {
                        const float4 w = *reinterpret_cast<const float4*>(tv.pf + p);
                        const float ex0 = sqrx - w.x, ey0 = sqry - w.y, ex1 = sqrx - w.z, ey1 = sqry - w.w;
                        if (p >= sp0) visit(p, nn_fma(ex0, ex0, ey0 * ey0));
                        visit(p + 1, nn_fma(ex1, ex1, ey1 * ey1));
                    } else if (p < sp1 && p >= sp0) {
                        const float2 v = tv.pf[p];
                        const float ex = sqrx - v.x, ey = sqry - v.y;
                        visit(p, nn_fma(ex, ex, ey * ey));
                    }
                }
            }
        }
    }
    // merge the lanes' filter states
#pragma unroll
    for (int d = G / 2; d >= 1; d >>= 1) {
        const float om1 = __uint_as_float(__shfl_xor_sync(ICP_FULL_MASK, __float_as_uint(f.m1), d));
        const float om2 = __uint_as_float(__shfl_xor_sync(ICP_FULL_MASK, __float_as_uint(f.m2), d));
        const int opos = (int)__shfl_xor_sync(ICP_FULL_MASK, (uint32_t)f.pos, d);
        f.m2 = fminf(fminf(f.m2, om2), fmaxf(f.m1, om1));
        f.pos = (om1 < f.m1) ? opos : f.pos;
        f.m1 = fminf(f.m1, om1);
    }
    int pos = -1;
    if (active) {
        const float thr = nn_filter_band(f.m1, Q.a);
        if (walk && f.m2 > thr && f.m1 < 3.0e38f) pos = f.pos;          // a single candidate inside the band
        else if (gl == 0) pos = walk ? nn_exact_pass(tv, Q, qx, qy, thr).pos : nn_ring_search(tv, qx, qy).pos;
    }
    return __shfl_sync(ICP_FULL_MASK, pos, gbase);
}

constexpr uint32_t kNnMaxPoints = 1u << 31;   // positions and original indices travel as int

// ---------------------------------------------------------------------------------------------
// Per-pair grids (shared memory / L2 slab: icp_lean.cuh, icp_stream.cuh) and 3-D shared targets:
// the same disc (ball) walk seeded with the match the source point had in the previous outer
// iteration.  ICP moves the scan by less and less, so that match is usually still the nearest point
// or next to it, and the disc covers one or two cells where the ring search always opens 3x3(x3).
// No binary32 copies here: the few candidates are ranked directly on their exact binary64 d2 (lowest
// original index on ties).
//
// bound (3-D): a skipped point lies in a slab outside [za, zb] (|dz|/h > Uc), or in a row of slab zz
// outside [ya, yb] (dy^2 + dz^2 > h^2 (ry^2 + zdc^2) >= h^2 Uc2), or in a column outside [xa, xb]
// (dx^2 + dy^2 + dz^2 > h^2 (wc^2 + ydc^2 + zdc^2) >= h^2 Uc2): its computed d2 exceeds the seed's.
// Array-of-structs copy of a sorted target (large per-pair grids in the L2 slab, icp_stream.cuh):
// one candidate = one 32-byte sector instead of three or four scattered 8-byte loads.
struct GridRec {
    double x, y, z;
    uint32_t idx, pad;
};
constexpr int kNnGroupDepth = 4;   // records per lane and step of the group search

ICP_DEV bool nn_better(double d2, int idx, double bd2, int bidx) { return d2 < bd2 || (d2 == bd2 && idx < bidx); }

// One record = two 16-byte loads of one 32-byte sector.  (One 256-bit load — sm_100's LDG.E.256 — was
// measured: 3.68k against 3.83k pairs/s on config 4.)
ICP_DEV GridRec nn_load_record(const GridRec* rec, int p) { return rec[p]; }

// Candidates [p0, p1) against `best`, in ascending position, on records: four at a time with their
// eight 16-byte loads issued before any of them is used — the records sit in L2 (~250 cycles away)
// and a CTA of the streaming kernel has only 16 warps to hide that with.
template <int DIM, typename IdxT, typename CellT>
ICP_DEV void nn_scan_records(const GridRec* rec, int p0, int p1, double qx, double qy, double qz,
                             NnResult<DIM, IdxT, CellT>& best)
{
    auto one = [&](const GridRec& c, int p) {
        const double dx = qx - c.x, dy = qy - c.y;
        double d2 = dx * dx + dy * dy;
        if (DIM == 3) {
            const double dz = qz - c.z;
            d2 = d2 + dz * dz;
        }
        if (d2 <= best.d2) {
            const int id = (int)c.idx;
            if (d2 < best.d2 || id < best.idx) { best.d2 = d2; best.idx = id; best.pos = p; }
        }
    };
    int p = p0;
    for (; p + 4 <= p1; p += 4) {
        const GridRec c0 = nn_load_record(rec, p), c1 = nn_load_record(rec, p + 1), c2 = nn_load_record(rec, p + 2), c3 = nn_load_record(rec, p + 3);
        one(c0, p); one(c1, p + 1); one(c2, p + 2); one(c3, p + 3);
    }
    for (; p < p1; ++p) {
        const GridRec c = nn_load_record(rec, p);
        one(c, p);
    }
}

// rec != nullptr: candidates are read from the array-of-structs copy (nn_scan_records)
// own_base / own_x >= 0 (3-D, records): the query's own cell is row base `own_base`, column `own_x`;
// all of its points are ranked BEFORE the ball is formed and the walk leaves that cell out.  For a
// first iteration (no previous match) this turns a ball through a mere member of the cell into
// one through the cell's best point.
template <int DIM, typename IdxT, typename CellT>
ICP_DEV NnResult<DIM, IdxT, CellT> grid_nn_search_seeded(const GridView<DIM, IdxT, CellT>& gv, double qx, double qy,
                                                         double qz, int seed_pos, const GridRec* rec = nullptr,
                                                         int own_base = -1, int own_x = -1)
{
    const GridGeom& g = gv.g;
    NnResult<DIM, IdxT, CellT> best;
    if (rec) {
        const GridRec s = rec[seed_pos];
        const double dx = qx - s.x, dy = qy - s.y;
        best.d2 = dx * dx + dy * dy;
        if (DIM == 3) {
            const double dz = qz - s.z;
            best.d2 = best.d2 + dz * dz;
        }
        best.pos = seed_pos;
        best.idx = (int)s.idx;
        if (DIM == 3 && own_x >= 0)
            nn_scan_records(rec, (int)gv.cstart[own_base + own_x], (int)gv.cstart[own_base + own_x + 1], qx, qy, qz, best);
    } else {
        const double dx = qx - gv.x[seed_pos], dy = qy - gv.y[seed_pos];
        best.d2 = dx * dx + dy * dy;
        if (DIM == 3) {
            const double dz = qz - gv.z[seed_pos];
            best.d2 = best.d2 + dz * dz;
        }
        best.pos = seed_pos;
        best.idx = (int)gv.idx[seed_pos];
    }
    NnQuery Q;
    Q.tx = nn_cell_coord(qx, g.ox, g.inv_h);
    Q.ty = nn_cell_coord(qy, g.oy, g.inv_h);
    Q.txf = (float)Q.tx; Q.tyf = (float)Q.ty;
    const float tzf = (DIM == 3) ? (float)nn_cell_coord(qz, g.oz, g.inv_h) : 0.0f;
    const float ihf = (float)g.inv_h;
    Q.Uc2 = ((float)best.d2 * ihf) * ihf * 1.000001f + 1e-30f;      // as in nn_query_setup
    if (!(fabsf(Q.txf) + fabsf(Q.tyf) + fabsf(tzf) <= 1e9f) || !(Q.Uc2 <= 1e12f) || !(best.d2 >= 0.0))
        return grid_nn_search(gv, qx, qy, qz);
    const float Uc = sqrtf(Q.Uc2) * 1.000001f;
    Q.sx = (fabsf(Q.txf) + 1.0f) * 3.6e-7f;
    Q.sy = (fabsf(Q.tyf) + 1.0f) * 3.6e-7f;
    if (DIM == 2) {   // one to three rows: plain loops are the cheapest
        const int ya = nn_floor_clamp(Q.tyf - Uc - Q.sy, g.gy - 1);     // bound (rows)
        const int yb2 = nn_floor_clamp(Q.tyf + Uc + Q.sy, g.gy - 1);
        for (int y2 = ya; y2 <= yb2; ++y2) {
            int xa, xb;
            if (!nn_row_span(Q, y2, g.gx, &xa, &xb)) continue;          // bound (columns)
            if (rec) nn_scan_records(rec, (int)gv.cstart[y2 * g.gx + xa], (int)gv.cstart[y2 * g.gx + xb + 1], qx, qy, qz, best);
            else grid_scan_cells(gv, y2 * g.gx, xa, xb, qx, qy, qz, best);
        }
        return best;
    }
    // 3-D walk shaped like nn_walk: each trip advances to the next non-empty span (slab zz, row yy),
    // then scans its candidates, so that the lanes of a warp run their candidate loops together.
    const float sz = (DIM == 3) ? (fabsf(tzf) + 1.0f) * 3.6e-7f : 0.0f;
    float Uc2 = Q.Uc2;
    int zz = 0, zb = 0;
    if (DIM == 3) {
        zz = nn_floor_clamp(tzf - Uc - sz, g.gz - 1);                   // bound (slabs)
        zb = nn_floor_clamp(tzf + Uc + sz, g.gz - 1);
    }
    int yy = 1, yb = 0;          // rows of the current slab still to visit
    bool slab_open = false;
    int pend0 = 0, pend1 = 0;    // second half of a span that was split around the own cell
    for (;;) {
        int p0 = pend0, p1 = pend1;
        pend0 = 0; pend1 = 0;
        while (p1 == p0 && zz <= zb) {
            if (!slab_open) {
                float rem2 = Uc2, ry = Uc;
                if (DIM == 3) {
                    const float zf = (float)zz;
                    const float zdc = fmaxf(fmaxf(zf - tzf, tzf - (zf + 1.0f)) - sz, 0.0f) * 0.999999f;
                    rem2 = Uc2 - zdc * zdc;
                    if (rem2 < 0.0f) { ++zz; continue; }
                    ry = sqrtf(rem2) * 1.000001f;
                }
                Q.Uc2 = rem2;    // nn_row_span tests wc^2 + ydc^2 against Uc2 - zdc^2
                yy = nn_floor_clamp(Q.tyf - ry - Q.sy, g.gy - 1);       // bound (rows)
                yb = nn_floor_clamp(Q.tyf + ry + Q.sy, g.gy - 1);
                slab_open = true;
            }
            if (yy > yb) { slab_open = false; ++zz; continue; }
            int xa, xb;
            if (nn_row_span(Q, yy, g.gx, &xa, &xb)) {                    // bound (columns)
                const int base = (zz * g.gy + yy) * g.gx;
                p0 = (int)gv.cstart[base + xa];
                p1 = (int)gv.cstart[base + xb + 1];
                if (base == own_base && xa <= own_x && own_x <= xb) {   // the own cell has been ranked
                    pend0 = (int)gv.cstart[base + own_x + 1];
                    pend1 = p1;
                    p1 = (int)gv.cstart[base + own_x];
                    if (p1 == p0) { p0 = pend0; p1 = pend1; pend0 = 0; pend1 = 0; }
                }
            }
            ++yy;
        }
        if (p1 == p0) break;
        const double before = best.d2;
        if (rec) {
            nn_scan_records(rec, p0, p1, qx, qy, qz, best);
        } else {
            for (int p = p0; p < p1; ++p) {
                const double dx = qx - gv.x[p], dy = qy - gv.y[p];
                double d2 = dx * dx + dy * dy;
                if (DIM == 3) {
                    const double dz = qz - gv.z[p];
                    d2 = d2 + dz * dz;
                }
                if (d2 <= best.d2) {
                    const int id = (int)gv.idx[p];
                    if (d2 < best.d2 || id < best.idx) { best.d2 = d2; best.idx = id; best.pos = p; }
                }
            }
        }
        // A closer point shrinks the ball for the slabs and rows still ahead: every bound above holds
        // for any point's d2 in place of the seed's (a skipped point's computed d2 exceeds it).  The
        // seed of a first iteration is only the first point of the query's cell — a 3-D cloud's
        // cells hold 50-150 points, so the first span usually brings the ball down to a cell or two.
        if (best.d2 < before) {
            const float u2 = ((float)best.d2 * ihf) * ihf * 1.000001f + 1e-30f;
            if (u2 < Uc2) {
                Uc2 = u2;
                const float Ucs = sqrtf(Uc2) * 1.000001f;
                if (DIM == 3) {
                    const int zb2 = nn_floor_clamp(tzf + Ucs + sz, g.gz - 1);                 // bound (slabs)
                    zb = zb2 < zb ? zb2 : zb;
                }
                if (slab_open) {    // the rows of the current slab that are still ahead
                    float rem2 = Uc2, ry = Ucs;
                    if (DIM == 3) {
                        const float zf = (float)zz;
                        const float zdc = fmaxf(fmaxf(zf - tzf, tzf - (zf + 1.0f)) - sz, 0.0f) * 0.999999f;
                        rem2 = Uc2 - zdc * zdc;
                        if (rem2 < 0.0f) { slab_open = false; ++zz; }
                        else ry = sqrtf(rem2) * 1.000001f;
                    }
                    if (slab_open) {
                        Q.Uc2 = rem2;
                        const int yb2 = nn_floor_clamp(Q.tyf + ry + Q.sy, g.gy - 1);          // bound (rows)
                        yb = yb2 < yb ? yb2 : yb;
                    }
                }
            }
        }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// The same seeded ball walk by a GROUP OF LANES for one query (large per-pair grids, icp_stream.cuh):
// 3-D LiDAR clouds put 50-150 points into a cell near the sensor, so a query meets a few LONG
// contiguous runs of candidates.  One thread per query walks them at ~7 of 32 lanes with three
// scattered 8-byte loads per candidate; here the G lanes of a group
//   - each take one (slab, row) of the ball's cover and form its span with the bounds above,
//   - then scan every non-empty span together, G consecutive 32-byte records per step (one
//     coalesced access), each lane keeping its own best on the exact binary64 d2,
//   - and finish with a shuffle arg-min on (d2, original index): the minimum of a totally ordered
//     set does not depend on the order it is reduced in, so the result is the per-thread walk's.
// (One query per whole warp was measured first: 1.57k pairs/s on config 4 against 2.76k for one
// thread per query — a warp then has a single dependent chain of L2 round trips in flight.)
// Records: the target sorted by cell as array-of-structs {x, y, z, idx} (GridRec), written by the
// grid build next to the SoA arrays.  All lanes pass the same arguments and return the same result.
// G lanes (a power of two, 4..32) work on one query; a warp holds 32 / G queries at a time and stays
// converged: every step each group scans G consecutive records of its current span, groups that have
// run out of candidates idle until the slowest one is done.  Call with all 32 lanes; the lanes of a
// group pass the same query; `active` false for a group without a query (it only takes part in the
// shuffles).  `ok` false on return (group-uniform) when the query is outside the walk's range: the
// caller then runs grid_nn_search for it.
template <int DIM, int G, typename IdxT, typename CellT>
ICP_DEV NnResult<DIM, IdxT, CellT> grid_nn_search_seeded_group(const GridView<DIM, IdxT, CellT>& gv, const GridRec* rec,
                                                               double qx, double qy, double qz, int seed_pos, bool active,
                                                               bool* ok)
{
    const GridGeom& g = gv.g;
    const int lane = lane_id();
    const int gl = lane & (G - 1);          // lane within the group
    const int gbase = lane & ~(G - 1);      // first lane of the group
    NnResult<DIM, IdxT, CellT> best;
    best.d2 = 1.0 / 0.0; best.idx = 0x7fffffff; best.pos = -1;
    NnQuery Q;
    Q.tx = 0.; Q.ty = 0.; Q.txf = 0.f; Q.tyf = 0.f; Q.sx = 0.f; Q.sy = 0.f; Q.Uc2 = 0.f;
    float tzf = 0.f, Uc2 = 0.f, sz = 0.f;
    int za = 0, zb = -1, ya = 0, ny = 1;
    bool good = false;
    if (active) {
        const GridRec s = rec[seed_pos];
        const double dx = qx - s.x, dy = qy - s.y;
        best.d2 = dx * dx + dy * dy;
        if (DIM == 3) {
            const double dz = qz - s.z;
            best.d2 = best.d2 + dz * dz;
        }
        best.pos = seed_pos;
        best.idx = (int)s.idx;
        Q.tx = nn_cell_coord(qx, g.ox, g.inv_h);
        Q.ty = nn_cell_coord(qy, g.oy, g.inv_h);
        Q.txf = (float)Q.tx; Q.tyf = (float)Q.ty;
        tzf = (DIM == 3) ? (float)nn_cell_coord(qz, g.oz, g.inv_h) : 0.0f;
        const float ihf = (float)g.inv_h;
        Uc2 = ((float)best.d2 * ihf) * ihf * 1.000001f + 1e-30f;      // as in nn_query_setup
        good = (fabsf(Q.txf) + fabsf(Q.tyf) + fabsf(tzf) <= 1e9f) && (Uc2 <= 1e12f) && (best.d2 >= 0.0);
        if (good) {
            const float Uc = sqrtf(Uc2) * 1.000001f;
            Q.sx = (fabsf(Q.txf) + 1.0f) * 3.6e-7f;
            Q.sy = (fabsf(Q.tyf) + 1.0f) * 3.6e-7f;
            sz = (DIM == 3) ? (fabsf(tzf) + 1.0f) * 3.6e-7f : 0.0f;
            za = 0; zb = 0;
            if (DIM == 3) {
                za = nn_floor_clamp(tzf - Uc - sz, g.gz - 1);                   // bound (slabs)
                zb = nn_floor_clamp(tzf + Uc + sz, g.gz - 1);
            }
            // rows of the widest slab: a superset of every slab's rows (each lane re-derives its slab's bound)
            ya = nn_floor_clamp(Q.tyf - Uc - Q.sy, g.gy - 1);                   // bound (rows)
            const int yb = nn_floor_clamp(Q.tyf + Uc + Q.sy, g.gy - 1);
            ny = yb - ya + 1;
        }
    }
    *ok = good || !active;
    const int nrows = good ? (zb - za + 1) * ny : 0;
    const int max_rows = __reduce_max_sync(ICP_FULL_MASK, nrows);
    for (int r0 = 0; r0 < max_rows; r0 += G) {
        // ---- lane gl of the group: span of row r0 + gl
        int p0 = 0, p1 = 0;
        const int r = r0 + gl;
        if (r < nrows) {
            const int zz = za + r / ny, yy = ya + r % ny;
            float rem2 = Uc2;
            bool live = true;
            if (DIM == 3) {
                const float zf = (float)zz;
                const float zdc = fmaxf(fmaxf(zf - tzf, tzf - (zf + 1.0f)) - sz, 0.0f) * 0.999999f;
                rem2 = Uc2 - zdc * zdc;
                live = !(rem2 < 0.0f);
            }
            if (live) {
                Q.Uc2 = rem2;    // nn_row_span tests wc^2 + ydc^2 against Uc2 - zdc^2 (bound (3-D))
                int xa, xb;
                if (nn_row_span(Q, yy, g.gx, &xa, &xb)) {                // bound (columns)
                    const int base = (zz * g.gy + yy) * g.gx;
                    p0 = (int)gv.cstart[base + xa];
                    p1 = (int)gv.cstart[base + xb + 1];
                }
            }
        }
        // ---- the group's non-empty spans, G records per step
        uint32_t todo = (__ballot_sync(ICP_FULL_MASK, p1 > p0) >> gbase) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
        int off = 0;
        while (__any_sync(ICP_FULL_MASK, todo != 0u)) {
            const int src = todo ? gbase + __ffs((int)todo) - 1 : lane;
            const int s0 = __shfl_sync(ICP_FULL_MASK, p0, src), s1 = __shfl_sync(ICP_FULL_MASK, p1, src);
            if (todo) {
                // kNnGroupDepth records per lane and step, all loads issued before the first use: the
                // records sit in L2 and the step is otherwise one dependent round trip
                const int pb = s0 + off + gl;
                GridRec c[kNnGroupDepth];
#pragma unroll
                for (int k = 0; k < kNnGroupDepth; ++k) {
                    const int p = pb + k * G;
                    c[k] = nn_load_record(rec, p < s1 ? p : s0);      // a clamped (harmless) load keeps the batch branch-free
                }
#pragma unroll
                for (int k = 0; k < kNnGroupDepth; ++k) {
                    const int p = pb + k * G;
                    const double dx = qx - c[k].x, dy = qy - c[k].y;
                    double d2 = dx * dx + dy * dy;
                    if (DIM == 3) {
                        const double dz = qz - c[k].z;
                        d2 = d2 + dz * dz;
                    }
                    if (p < s1 && nn_better(d2, (int)c[k].idx, best.d2, best.idx)) { best.d2 = d2; best.idx = (int)c[k].idx; best.pos = p; }
                }
                off += kNnGroupDepth * G;
                if (s0 + off >= s1) { todo &= todo - 1u; off = 0; }
            }
        }
    }
    // ---- arg-min over the group's lanes on (d2, original index)
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1) {
        const double od2 = shfl_xor_f64(best.d2, m);
        const int oidx = (int)__shfl_xor_sync(ICP_FULL_MASK, (uint32_t)best.idx, m);
        const int opos = (int)__shfl_xor_sync(ICP_FULL_MASK, (uint32_t)best.pos, m);
        if (nn_better(od2, oidx, best.d2, best.idx)) { best.d2 = od2; best.idx = oidx; best.pos = opos; }
    }
    return best;
}

// Number of (slab, row) pairs the seeded ball walk of a query would visit — the cost measure the
// streaming kernel uses to hand queries with a far seed (a source point without a counterpart: its
// previous match is decimetres away and the ball covers hundreds of mostly empty rows) to a whole
// warp instead of stalling the 31 other lanes of their own warp.  0: outside the walk's range.
template <int DIM, typename IdxT, typename CellT>
ICP_DEV int grid_ball_rows(const GridView<DIM, IdxT, CellT>& gv, const GridRec* rec, double qx, double qy, double qz,
                           int seed_pos)
{
    const GridGeom& g = gv.g;
    const GridRec s = rec[seed_pos];
    const double dx = qx - s.x, dy = qy - s.y;
    double d2 = dx * dx + dy * dy;
    if (DIM == 3) {
        const double dz = qz - s.z;
        d2 = d2 + dz * dz;
    }
    const float txf = (float)nn_cell_coord(qx, g.ox, g.inv_h), tyf = (float)nn_cell_coord(qy, g.oy, g.inv_h);
    const float tzf = (DIM == 3) ? (float)nn_cell_coord(qz, g.oz, g.inv_h) : 0.0f;
    const float ihf = (float)g.inv_h;
    const float Uc2 = ((float)d2 * ihf) * ihf * 1.000001f + 1e-30f;
    if (!(fabsf(txf) + fabsf(tyf) + fabsf(tzf) <= 1e9f) || !(Uc2 <= 1e12f) || !(d2 >= 0.0)) return 0;
    const float Uc = sqrtf(Uc2) * 1.000001f;
    const int ny = nn_floor_clamp(tyf + Uc + 1.0f, g.gy - 1) - nn_floor_clamp(tyf - Uc - 1.0f, g.gy - 1) + 1;
    const int nz = (DIM == 3) ? nn_floor_clamp(tzf + Uc + 1.0f, g.gz - 1) - nn_floor_clamp(tzf - Uc - 1.0f, g.gz - 1) + 1 : 1;
    return ny * nz;
}

// ---------------------------------------------------------------------------------------------
// Binary32 filter for the 3-D record walk (large per-pair grids, icp_stream.cuh; config 4).
// The exact walk above spends its time on 32-byte records and binary64 arithmetic at ~9 of 32
// lanes.  Here every candidate is first ranked on a 16-byte record {x, y, z in CELL UNITS from the
// grid's origin, binary32; original index} — half the bytes, a third of the arithmetic — exactly as
// target_nn_search2 does on a shared 2-D target, with the same argument:
//   error: the stored coordinate and the query's coordinate are each rounded once to binary32
//          (|.| <= A := max|T(q)| + G, G the largest grid dimension) and their difference once more:
//          each of ex, ey, ez is within a := 2^-23 (A + 2)(1 + 1e-2) of the true (q - p)/h;
//   bound (filter, 3-D): sqrt(d32) is within sqrt(3) a + 2^-22 sqrt(d32) of the true distance D/h,
//          so the exact winner w has sqrt(d32(w)) <= sqrt(m1)(1 + 2^-20) + 4a for the binary32
//          minimum m1: candidates above the band (nn_filter_band3, 6a) cannot win;
//   a single candidate inside the band IS the winner; two or more (near-ties at the 1e-5-cell
//   level, duplicates) send the query to the exact binary64 walk above, unchanged.
// The cover (slabs, rows, columns) is the exact walk's, formed from the seed's exact d2; the own
// cell first and the shrinking ball work on the filter's running minimum (a candidate with
// binary32 distance sqrt(m1) is at most (sqrt(m1)(1 + 2^-22) + sqrt(3) a) h away).
struct GridRecF {
    float x, y, z;
    uint32_t idx;
};

ICP_DEV float nn_filter_band3(float m1, float a)
{
    const float t = sqrtf(m1) * 1.000004f + 6.0f * a;
    return t * t * 1.000001f + 1e-30f;
}

// Candidates [p0, p1) through the filter, eight 16-byte loads in flight.
ICP_DEV void nn_scan_records_f32(const GridRecF* recf, int p0, int p1, float qxf, float qyf, float qzf, NnFilterState& f)
{
    auto one = [&](const float4& c, int p) {
        const float ex = qxf - c.x, ey = qyf - c.y, ez = qzf - c.z;
        const float d32 = nn_fma(ex, ex, nn_fma(ey, ey, ez * ez));
        f.m2 = fminf(f.m2, fmaxf(d32, f.m1));
        f.pos = (d32 < f.m1) ? p : f.pos;
        f.m1 = fminf(f.m1, d32);
    };
    const float4* r4 = reinterpret_cast<const float4*>(recf);
    int p = p0;
    for (; p + 8 <= p1; p += 8) {
        const float4 c0 = r4[p], c1 = r4[p + 1], c2 = r4[p + 2], c3 = r4[p + 3];
        const float4 c4 = r4[p + 4], c5 = r4[p + 5], c6 = r4[p + 6], c7 = r4[p + 7];
        one(c0, p); one(c1, p + 1); one(c2, p + 2); one(c3, p + 3);
        one(c4, p + 4); one(c5, p + 5); one(c6, p + 6); one(c7, p + 7);
    }
    if (p + 4 <= p1) {
        const float4 c0 = r4[p], c1 = r4[p + 1], c2 = r4[p + 2], c3 = r4[p + 3];
        one(c0, p); one(c1, p + 1); one(c2, p + 2); one(c3, p + 3);
        p += 4;
    }
    for (; p < p1; ++p) {
        const float4 c = r4[p];
        one(c, p);
    }
}

template <typename IdxT, typename CellT>
ICP_DEV NnResult<3, IdxT, CellT> grid_nn_search_seeded_f32(const GridView<3, IdxT, CellT>& gv, const GridRec* rec,
                                                           const GridRecF* recf, double qx, double qy, double qz,
                                                           int seed_pos, int own_base = -1, int own_x = -1)
{
    const GridGeom& g = gv.g;
    const GridRec s = rec[seed_pos];
    const double sdx = qx - s.x, sdy = qy - s.y, sdz = qz - s.z;
    const double U2 = (sdx * sdx + sdy * sdy) + sdz * sdz;
    NnQuery Q;
    Q.tx = nn_cell_coord(qx, g.ox, g.inv_h);
    Q.ty = nn_cell_coord(qy, g.oy, g.inv_h);
    Q.txf = (float)Q.tx; Q.tyf = (float)Q.ty;
    const float tzf = (float)nn_cell_coord(qz, g.oz, g.inv_h);
    const float ihf = (float)g.inv_h;
    float Uc2 = ((float)U2 * ihf) * ihf * 1.000001f + 1e-30f;      // as in nn_query_setup
    const float amax = fmaxf(fmaxf(fabsf(Q.txf), fabsf(Q.tyf)), fabsf(tzf));
    const int gmax = g.gx > g.gy ? (g.gx > g.gz ? g.gx : g.gz) : (g.gy > g.gz ? g.gy : g.gz);
    if (!(amax <= 1e6f) || !(Uc2 <= 1e12f) || !(U2 >= 0.0))
        return grid_nn_search_seeded(gv, qx, qy, qz, seed_pos, rec, own_base, own_x);
    const float a = (amax + (float)gmax + 2.0f) * 1.21e-7f;    // 2^-23 (A + 2)(1 + 1e-2)
    Q.sx = (fabsf(Q.txf) + 1.0f) * 3.6e-7f;
    Q.sy = (fabsf(Q.tyf) + 1.0f) * 3.6e-7f;
    const float sz = (fabsf(tzf) + 1.0f) * 3.6e-7f;
    NnFilterState f;
    f.m1 = 3.0e38f; f.m2 = 3.0e38f; f.pos = seed_pos;
    // a closer candidate shrinks the ball: its true distance is at most (sqrt(m1)(1 + 2^-22) + sqrt(3) a) h
    auto tighten = [&]() {
        const float uc = sqrtf(f.m1) * 1.000001f + 2.0f * a;
        const float u2 = uc * uc * 1.000002f;
        const bool smaller = u2 < Uc2;
        Uc2 = smaller ? u2 : Uc2;
        return smaller;
    };
    if (own_x >= 0) {
        nn_scan_records_f32(recf, (int)gv.cstart[own_base + own_x], (int)gv.cstart[own_base + own_x + 1], Q.txf, Q.tyf, tzf, f);
        tighten();
    }
    float Uc = sqrtf(Uc2) * 1.000001f;
    int zz = nn_floor_clamp(tzf - Uc - sz, g.gz - 1);                   // bound (slabs)
    int zb = nn_floor_clamp(tzf + Uc + sz, g.gz - 1);
    int yy = 1, yb = 0;          // rows of the current slab still to visit
    bool slab_open = false;
    int pend0 = 0, pend1 = 0;    // second half of a span that was split around the own cell
    Q.Uc2 = Uc2;
    for (;;) {
        int p0 = pend0, p1 = pend1;
        pend0 = 0; pend1 = 0;
        while (p1 == p0 && zz <= zb) {
            if (!slab_open) {
                const float zf = (float)zz;
                const float zdc = fmaxf(fmaxf(zf - tzf, tzf - (zf + 1.0f)) - sz, 0.0f) * 0.999999f;
                const float rem2 = Uc2 - zdc * zdc;
                if (rem2 < 0.0f) { ++zz; continue; }
                const float ry = sqrtf(rem2) * 1.000001f;
                Q.Uc2 = rem2;    // nn_row_span tests wc^2 + ydc^2 against Uc2 - zdc^2
                yy = nn_floor_clamp(Q.tyf - ry - Q.sy, g.gy - 1);       // bound (rows)
                yb = nn_floor_clamp(Q.tyf + ry + Q.sy, g.gy - 1);
                slab_open = true;
            }
            if (yy > yb) { slab_open = false; ++zz; continue; }
            int xa, xb;
            if (nn_row_span(Q, yy, g.gx, &xa, &xb)) {                    // bound (columns)
                const int base = (zz * g.gy + yy) * g.gx;
                p0 = (int)gv.cstart[base + xa];
                p1 = (int)gv.cstart[base + xb + 1];
                if (base == own_base && xa <= own_x && own_x <= xb) {   // the own cell has been ranked
                    pend0 = (int)gv.cstart[base + own_x + 1];
                    pend1 = p1;
                    p1 = (int)gv.cstart[base + own_x];
                    if (p1 == p0) { p0 = pend0; p1 = pend1; pend0 = 0; pend1 = 0; }
                }
            }
            ++yy;
        }
        if (p1 == p0) break;
        const float before = f.m1;
        nn_scan_records_f32(recf, p0, p1, Q.txf, Q.tyf, tzf, f);
        if (f.m1 < before && tighten()) {
            Uc = sqrtf(Uc2) * 1.000001f;
            const int zb2 = nn_floor_clamp(tzf + Uc + sz, g.gz - 1);                 // bound (slabs)
            zb = zb2 < zb ? zb2 : zb;
            if (slab_open) {    // the rows of the current slab that are still ahead
                const float zf = (float)zz;
                const float zdc = fmaxf(fmaxf(zf - tzf, tzf - (zf + 1.0f)) - sz, 0.0f) * 0.999999f;
                const float rem2 = Uc2 - zdc * zdc;
                if (rem2 < 0.0f) {
                    slab_open = false; ++zz;
                } else {
                    Q.Uc2 = rem2;
                    const int yb2 = nn_floor_clamp(Q.tyf + sqrtf(rem2) * 1.000001f + Q.sy, g.gy - 1);   // bound (rows)
                    yb = yb2 < yb ? yb2 : yb;
                }
            }
        }
    }
    const float thr = nn_filter_band3(f.m1, a);
    if (f.m2 > thr && f.m1 < 3.0e38f) {   // a single candidate inside the band: the winner
        NnResult<3, IdxT, CellT> r;
        r.d2 = 0.0; r.pos = f.pos; r.idx = (int)recf[f.pos].idx;
        return r;
    }
    return grid_nn_search_seeded(gv, qx, qy, qz, seed_pos, rec, own_base, own_x);
}

// Search of a per-pair grid with whatever seed is at hand: the previous match (seed_pos >= 0), else
// the first point of the query's own cell, else (empty cell) the ring search.
template <int DIM, typename IdxT, typename CellT>
ICP_DEV NnResult<DIM, IdxT, CellT> grid_nn_search_auto(const GridView<DIM, IdxT, CellT>& gv, double qx, double qy,
                                                       double qz, int seed_pos, const GridRec* rec = nullptr,
                                                       const GridRecF* recf = nullptr)
{
    if (seed_pos < 0) {
        const GridGeom& g = gv.g;
        const int cx = grid_coord(qx, g.ox, g.inv_h, g.gx);
        const int cy = grid_coord(qy, g.oy, g.inv_h, g.gy);
        const int cz = (DIM == 3) ? grid_coord(qz, g.oz, g.inv_h, g.gz) : 0;
        const int c = (cz * g.gy + cy) * g.gx + cx;
        const int p0 = (int)gv.cstart[c], p1 = (int)gv.cstart[c + 1];
        if (p1 > p0) {
            // 3-D records: rank the whole own cell first and leave it out of the walk.  (The best point
            // of the own cell as a mere seed, the cell then scanned again by the walk, was measured
            // before the walk could shrink its ball: 2.80k against 2.98k pairs/s.)
            if constexpr (DIM == 3) {
                if (rec && recf) return grid_nn_search_seeded_f32(gv, rec, recf, qx, qy, qz, p0, c - cx, cx);
                if (rec) return grid_nn_search_seeded(gv, qx, qy, qz, p0, rec, c - cx, cx);
            }
            seed_pos = p0;
        }
    }
    if constexpr (DIM == 3) {
        if (seed_pos >= 0 && rec && recf) return grid_nn_search_seeded_f32(gv, rec, recf, qx, qy, qz, seed_pos);
    }
    return seed_pos >= 0 ? grid_nn_search_seeded(gv, qx, qy, qz, seed_pos, rec) : grid_nn_search(gv, qx, qy, qz);
}

// seed_pos >= 0: a sorted position of the target (the previous outer iteration's match); -1: none
template <int DIM>
ICP_DEV NnResult<DIM, uint32_t, uint32_t> target_nn_search(const TargetView& tv, double qx, double qy, double qz,
                                                          int seed_pos = -1, bool need_d2 = true)
{
    if constexpr (DIM == 2) {
        if (tv.pf != nullptr) return target_nn_search2(tv, qx, qy, seed_pos, need_d2);
    }
    GridView<DIM, uint32_t, uint32_t> gv;
    gv.x = tv.x; gv.y = tv.y; gv.z = tv.z; gv.idx = tv.idx; gv.cstart = tv.cstart; gv.g = tv.g;
    if (seed_pos >= 0 && (uint32_t)seed_pos < tv.m) return grid_nn_search_seeded(gv, qx, qy, qz, seed_pos);
    return grid_nn_search(gv, qx, qy, qz);
}

}  // namespace icp
