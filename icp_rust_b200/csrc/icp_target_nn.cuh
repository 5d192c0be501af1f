// icp_target_nn.cuh — correspondence search on a large shared target (the `Icp2d<'a>` value kept on
// the device, src/lib.rs:91-102; KdTree::search per query, src/lib.rs:121).
//
// Contract (identical to icp_grid.cuh and the oracle's brute force): exact Euclidean 1-NN on the
// COMPUTED d2 = dx*dx + dy*dy (binary64, dx = q.x - p.x, no FMA), lowest original index on equal d2.
//
// 2-D targets carry, besides the sorted cell list of icp_grid.cuh,
//   seed[c]  u32     sorted position of a target point near cell c (jump flooding at build time);
//                    ANY point is a valid seed — its exact d2 bounds the search disc, nothing more
//   pf[p]    float2  binary32 copy of point p relative to the centre of its 32x32-cell block, so
//                    that a binary32 distance to a query carries an error of ~2^-22 (U + 40h)
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

// Origin of block b along one axis.  Build and search evaluate this same expression.
ICP_HD double nn_block_origin(double o, double h, int b)
{
    return o + (double)(b * kNnBlockCells + kNnBlockCells / 2) * h;
}

ICP_DEV float nn_fma(float a, float b, float c)
{
#ifdef ICP_SIMT_EMULATE
    return std::fmaf(a, b, c);
#else
    return __fmaf_rn(a, b, c);
#endif
}

#ifdef ICP_SIMT_EMULATE
// emulator-only work counters (tests/simt): queries, rows, spans, candidates of pass 1, pass-2 queries
struct NnCounters { unsigned long long q, rows, spans, cand, pass2, exact, pre; };
inline NnCounters& nn_counters() { static NnCounters c{}; return c; }
#define ICP_NN_COUNT(field, v) (nn_counters().field += (v))
#else
#define ICP_NN_COUNT(field, v) ((void)0)
#endif

struct NnFilterState {
    float m1, m2;   // smallest and second smallest binary32 d2 seen
    int pos;        // sorted position of the smallest
};

// Exact 1-NN on a 2-D shared target with seeds.  Requires tv.m > 0, tv.seed, tv.pf.
ICP_DEV NnResult<2, uint32_t, uint32_t> target_nn_search2(const TargetView& tv, double qx, double qy)
{
    const GridGeom& g = tv.g;
    const int cx = grid_coord(qx, g.ox, g.inv_h, g.gx);
    const int cy = grid_coord(qy, g.oy, g.inv_h, g.gy);
    NnResult<2, uint32_t, uint32_t> best;
    const uint32_t s = tv.seed[cy * g.gx + cx];
    if (s >= tv.m) {   // no seed (cannot happen after a complete flood): ring search
        GridView<2, uint32_t, uint32_t> gv;
        gv.x = tv.x; gv.y = tv.y; gv.z = tv.z; gv.idx = tv.idx; gv.cstart = tv.cstart; gv.g = g;
        return grid_nn_search(gv, qx, qy, 0.0);
    }
    {
        const double dx = qx - tv.x[s], dy = qy - tv.y[s];
        best.d2 = dx * dx + dy * dy;
        best.pos = (int)s;
    }
    const double U2 = best.d2;
    // binary32 upper bounds of U2 and sqrt(U2) (each factor 1.000001 covers one rounding of 2^-24)
    // (+1e-30: keeps the bounds valid where binary32 would go subnormal)
    const float U2f = (float)U2 * 1.000001f + 1e-30f;
    const float Uf = sqrtf(U2f) * 1.000001f;
    const float hf = (float)g.h;
    if (!(U2f <= 1e30f) || !(U2 >= 0.0)) {   // overflow of the filter's range / NaN: ring search
        GridView<2, uint32_t, uint32_t> gv;
        gv.x = tv.x; gv.y = tv.y; gv.z = tv.z; gv.idx = tv.idx; gv.cstart = tv.cstart; gv.g = g;
        return grid_nn_search(gv, qx, qy, 0.0);
    }
    // absolute slack that keeps q -+ radius on the safe side of its own rounding
    const double eps_abs = (fabs(qx) + fabs(qy)) * 1e-15;
    const double Ud = (double)Uf + eps_abs;
    // bound (rows): a point in a row above cell(fl(qy + Ud)) has p.y > fl(qy + Ud) because grid_coord
    // is monotone, so |q.y - p.y| > sqrt(U2)(1 + 5e-7) and its computed d2 exceeds U2.  Same below.
    const int ya = grid_coord(qy - Ud, g.oy, g.inv_h, g.gy);
    const int yb = grid_coord(qy + Ud, g.oy, g.inv_h, g.gy);
    // error bound of one binary32 coordinate difference: the stored point (|.| <= 17h) and the query
    // offset (|.| <= U + 18h) are rounded to 24 bits, and so is their difference (|.| <= U + 35h)
    const float a = (Uf + 40.0f * hf) * 2.4e-7f;   // 2^-22
    ICP_NN_COUNT(q, 1);
    NnFilterState f;
    f.m1 = 3.0e38f; f.m2 = 3.0e38f; f.pos = (int)s;

    // Walks the cells the disc reaches; `visit(p, d32)` sees every candidate's binary32 d2.
    auto walk = [&](auto&& visit) {
        for (int yy = ya; yy <= yb; ++yy) {
            // bound (columns): points of row yy have |q.y - p.y| >= yd (cell boundary up to `margin`,
            // icp_grid.cuh); outside [q.x - wd, q.x + wd] their computed d2 is
            // >= (wd^2 + yd^2)(1 - 2^-50) >= U2 (1 + 8e-7) > U2.  Monotone grid_coord again turns the
            // interval into a span of cells.
            double yd = 0.0;
            if (yy > cy) yd = (g.oy + (double)yy * g.h) - qy - g.margin;
            else if (yy < cy) yd = qy - (g.oy + (double)(yy + 1) * g.h) - g.margin;
            const float ydf = (float)fmax(yd, 0.0) * 0.999999f;
            const float w2 = U2f - ydf * ydf;
            if (w2 < 0.0f) continue;
            ICP_NN_COUNT(rows, 1);
            const double wd = (double)(sqrtf(w2) * 1.000001f) + eps_abs;
            const int xa = grid_coord(qx - wd, g.ox, g.inv_h, g.gx);
            const int xb = grid_coord(qx + wd, g.ox, g.inv_h, g.gx);
            const float qry = (float)(qy - nn_block_origin(g.oy, g.h, yy >> kNnBlockShift));
            const int row = yy * g.gx;
            for (int xs = xa; xs <= xb;) {
                const int xe_blk = xs | (kNnBlockCells - 1);
                const int xe = xe_blk < xb ? xe_blk : xb;
                const int p0 = (int)tv.cstart[row + xs];
                const int p1 = (int)tv.cstart[row + xe + 1];
                ICP_NN_COUNT(spans, 1);
                ICP_NN_COUNT(cand, p1 - p0);
                if (p1 > p0) {
                    const float qrx = (float)(qx - nn_block_origin(g.ox, g.h, xs >> kNnBlockShift));
                    for (int p = p0; p < p1; ++p) {
                        const float2 v = tv.pf[p];
                        const float ex = qrx - v.x, ey = qry - v.y;
                        visit(p, nn_fma(ex, ex, ey * ey));
                    }
                }
                xs = xe + 1;
            }
        }
    };

    walk([&](int p, float d32) {
        f.m2 = fminf(f.m2, fmaxf(d32, f.m1));
        f.pos = (d32 < f.m1) ? p : f.pos;
        f.m1 = fminf(f.m1, d32);
    });

    // bound (filter): sqrt(d32) is within 1.5a + 2^-22 sqrt(d32) of the true distance D.  The winner w
    // of the exact ranking has D(w) <= D(p)(1 + 2^-50) for every p, in particular for the binary32
    // minimum, hence sqrt(d32(w)) <= sqrt(m1)(1 + 2^-20) + 4a.  Candidates above `thr` cannot win.
    const float t = sqrtf(f.m1) * 1.000004f + 5.0f * a;
    const float thr = t * t * 1.000001f + 1e-30f;
    if (f.m2 > thr && f.m1 < 3.0e38f) {
        // a single candidate inside the band, and the exact winner is inside the band: it is the winner
        const int p = f.pos;
        const double dx = qx - tv.x[p], dy = qy - tv.y[p];
        best.d2 = dx * dx + dy * dy;
        best.pos = p;
        best.idx = (int)tv.idx[p];
        return best;
    }
    best.idx = (int)tv.idx[s];
    ICP_NN_COUNT(pass2, 1);
    walk([&](int p, float d32) {
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

template <int DIM>
ICP_DEV NnResult<DIM, uint32_t, uint32_t> target_nn_search(const TargetView& tv, double qx, double qy, double qz)
{
    if constexpr (DIM == 2) {
        if (tv.pf != nullptr) return target_nn_search2(tv, qx, qy);
    }
    GridView<DIM, uint32_t, uint32_t> gv;
    gv.x = tv.x; gv.y = tv.y; gv.z = tv.z; gv.idx = tv.idx; gv.cstart = tv.cstart; gv.g = tv.g;
    return grid_nn_search(gv, qx, qy, qz);
}

}  // namespace icp
