// icp_grid.cuh — uniform-grid / sorted cell-list nearest-neighbour search.
//
// Replaces nearest_neighbor::KdTree (external crate; call sites src/lib.rs:99,121,141,164).
// Contract (identical to the oracle's brute force): exact Euclidean 1-NN on
//     d2 = dx*dx + dy*dy (+ dz*dz),  dx = q.x - p.x,
// lowest ORIGINAL index wins on equal d2.  The grid only prunes; every pruning test is a
// proven lower bound on the computed d2 of all unvisited points (see "bound" below), so the
// result is bit-identical to scanning every target point.
#pragma once
#include "icp_cta.cuh"

namespace icp {

struct GridGeom {
    double ox, oy, oz;   // min corner of the target's bounding box
    double h, inv_h;     // cubic cell edge and fl(1/h)
    double margin;       // absolute slack covering rounding of cell boundaries
    int gx, gy, gz;      // cells per axis (gz == 1 in 2-D)
};

// Monotone non-decreasing in v (each step is a monotone IEEE operation), therefore points of
// cell c and points of cell c+1 are separated by a single threshold on v.
ICP_HD int grid_coord(double v, double o, double inv_h, int g)
{
    double t = (v - o) * inv_h;
    t = fmin(fmax(t, 0.0), (double)(g - 1));
    return (int)t;
}

// Chooses a cell edge so that the box fits `budget` cells.  Pure function of the bounding box:
// every thread computes the same geometry.  Geometry never changes results, only speed.
ICP_HD GridGeom grid_make_geom(double minx, double miny, double minz, double maxx, double maxy,
                               double maxz, int dim, int budget)
{
    GridGeom g;
    g.ox = minx; g.oy = miny; g.oz = (dim == 3) ? minz : 0.0;
    double ex = maxx - minx, ey = maxy - miny, ez = (dim == 3) ? (maxz - minz) : 0.0;
    double emax = fmax(ex, fmax(ey, ez));
    double h;
    if (!(emax > 0.0) || !(emax < 1.7e308)) {
        h = 1.0;
    } else {
        // measure of the occupied box, ignoring (nearly) flat directions
        double eps = emax * 1e-3;
        double vol = 1.0;
        int nd = 0;
        if (ex > eps) { vol *= ex; ++nd; }
        if (ey > eps) { vol *= ey; ++nd; }
        if (ez > eps) { vol *= ez; ++nd; }
        double per = vol / (double)budget;
        h = (nd == 1) ? per : (nd == 2 ? sqrt(per) : cbrt(per));
        h = fmax(h, emax / (double)budget);
        if (!(h > 0.0)) h = 1.0;
    }
    for (int it = 0; it < 64; ++it) {
        double cx = floor(ex / h) + 1.0, cy = floor(ey / h) + 1.0, cz = floor(ez / h) + 1.0;
        if (cx * cy * cz <= (double)budget && cx < 32000.0 && cy < 32000.0 && cz < 32000.0) {
            g.gx = (int)cx; g.gy = (int)cy; g.gz = (int)cz;
            break;
        }
        h *= 1.25;
        g.gx = 1; g.gy = 1; g.gz = 1;  // overwritten unless the loop runs out (h then covers the box)
    }
    g.h = h;
    g.inv_h = 1.0 / h;
    double span = h * (double)(g.gx > g.gy ? (g.gx > g.gz ? g.gx : g.gz) : (g.gy > g.gz ? g.gy : g.gz));
    g.margin = (fabs(g.ox) + fabs(g.oy) + fabs(g.oz) + 2.0 * span) * 3.6e-15;  // ~2^-48
    return g;
}

// Sorted cell list as the search sees it.  IdxT / CellT are u16 for per-pair grids staged in
// shared memory and u32 for large shared targets read from global memory.
template <int DIM, typename IdxT, typename CellT>
struct GridView {
    const double* x;
    const double* y;
    const double* z;        // DIM == 3 only
    const IdxT* idx;        // original index of the point stored at each sorted position
    const CellT* cstart;    // [ncells + 1] start of each cell in the sorted arrays
    GridGeom g;
};

template <int DIM, typename IdxT, typename CellT>
struct NnResult {
    double d2;
    int idx;   // original index
    int pos;   // sorted position (coordinates are x[pos], y[pos])
};

template <int DIM, typename IdxT, typename CellT>
ICP_DEV void grid_scan_cells(const GridView<DIM, IdxT, CellT>& gv, int row_base, int xa, int xb,
                             double qx, double qy, double qz, NnResult<DIM, IdxT, CellT>& best)
{
    int p0 = (int)gv.cstart[row_base + xa];
    int p1 = (int)gv.cstart[row_base + xb + 1];
    for (int p = p0; p < p1; ++p) {
        double dx = qx - gv.x[p];
        double dy = qy - gv.y[p];
        double d2 = dx * dx + dy * dy;
        if (DIM == 3) {
            double dz = qz - gv.z[p];
            d2 = d2 + dz * dz;
        }
        if (d2 <= best.d2) {
            int id = (int)gv.idx[p];
            if (d2 < best.d2 || id < best.idx) {
                best.d2 = d2;
                best.idx = id;
                best.pos = p;
            }
        }
    }
}

// One Chebyshev ring of one z-slab: rows ya..yb of the square [x0,x1]x[y0,y1] clipped to the
// grid; `full` scans whole rows (top/bottom slab of a 3-D shell), otherwise only the perimeter.
template <int DIM, typename IdxT, typename CellT>
ICP_DEV void grid_scan_ring_slab(const GridView<DIM, IdxT, CellT>& gv, int slab_base, int x0, int x1,
                                 int y0, int y1, bool full, double qx, double qy, double qz,
                                 NnResult<DIM, IdxT, CellT>& best)
{
    const int gx = gv.g.gx, gy = gv.g.gy;
    const int xa = x0 < 0 ? 0 : x0, xb = x1 > gx - 1 ? gx - 1 : x1;
    const int ya = y0 < 0 ? 0 : y0, yb = y1 > gy - 1 ? gy - 1 : y1;
    for (int yy = ya; yy <= yb; ++yy) {
        const int row_base = slab_base + yy * gx;
        if (full || yy == y0 || yy == y1) {
            grid_scan_cells(gv, row_base, xa, xb, qx, qy, qz, best);
        } else {
            if (x0 >= 0) grid_scan_cells(gv, row_base, x0, x0, qx, qy, qz, best);
            if (x1 <= gx - 1 && x1 != x0) grid_scan_cells(gv, row_base, x1, x1, qx, qy, qz, best);
        }
    }
}

// Exact 1-NN.  Requires at least one target point.
//
// bound: after rings 0..rho the visited block is the cell range [x0,x1]x[y0,y1](x[z0,z1]).  A
// point in an unvisited cell has a cell coordinate < x0 or > x1 on some axis; by monotonicity
// of grid_coord its coordinate then lies beyond the cell boundary B = o + c*h up to the
// rounding slack `margin`, so |q - p| >= L - margin with L the distance from q to the nearest
// face that still has cells behind it.  Its computed d2 is >= (L-margin)^2 (1 - 2^-50).  The
// search stops only when the best d2 is strictly below that, so no unvisited point can tie or
// beat it.
template <int DIM, typename IdxT, typename CellT>
ICP_DEV NnResult<DIM, IdxT, CellT> grid_nn_search(const GridView<DIM, IdxT, CellT>& gv, double qx,
                                                  double qy, double qz)
{
    const GridGeom& g = gv.g;
    const int cx = grid_coord(qx, g.ox, g.inv_h, g.gx);
    const int cy = grid_coord(qy, g.oy, g.inv_h, g.gy);
    const int cz = (DIM == 3) ? grid_coord(qz, g.oz, g.inv_h, g.gz) : 0;
    NnResult<DIM, IdxT, CellT> best;
    best.d2 = 1.0 / 0.0;
    best.idx = 0x7fffffff;
    best.pos = -1;
    const double inf = 1.0 / 0.0;
    // 2-D: the first step scans the whole 3x3 block (rings 0 and 1 together).  The lanes of a warp
    // almost never all stop after ring 0, so its separate bound test and row bookkeeping only
    // cost; the set of visited cells before each bound test is still a full block [x0,x1]x[y0,y1].
    for (int rho = (DIM == 2) ? 1 : 0;; ++rho) {
        const int x0 = cx - rho, x1 = cx + rho, y0 = cy - rho, y1 = cy + rho;
        if (DIM == 2) {
            grid_scan_ring_slab(gv, 0, x0, x1, y0, y1, rho == 1, qx, qy, qz, best);
        } else {
            const int z0 = cz - rho, z1 = cz + rho;
            const int za = z0 < 0 ? 0 : z0, zb = z1 > g.gz - 1 ? g.gz - 1 : z1;
            for (int zz = za; zz <= zb; ++zz)
                grid_scan_ring_slab(gv, zz * g.gy * g.gx, x0, x1, y0, y1, zz == z0 || zz == z1, qx,
                                    qy, qz, best);
        }
        double L = inf;
        if (x0 > 0) L = fmin(L, qx - (g.ox + (double)x0 * g.h));
        if (x1 < g.gx - 1) L = fmin(L, (g.ox + (double)(x1 + 1) * g.h) - qx);
        if (y0 > 0) L = fmin(L, qy - (g.oy + (double)y0 * g.h));
        if (y1 < g.gy - 1) L = fmin(L, (g.oy + (double)(y1 + 1) * g.h) - qy);
        if (DIM == 3) {
            const int z0 = cz - rho, z1 = cz + rho;
            if (z0 > 0) L = fmin(L, qz - (g.oz + (double)z0 * g.h));
            if (z1 < g.gz - 1) L = fmin(L, (g.oz + (double)(z1 + 1) * g.h) - qz);
        }
        if (L == inf) break;  // the whole grid has been visited
        const double Ls = L - g.margin;
        if (best.pos >= 0 && Ls > 0.0 && best.d2 < Ls * Ls * 0.999999999) break;
    }
    return best;
}

}  // namespace icp
