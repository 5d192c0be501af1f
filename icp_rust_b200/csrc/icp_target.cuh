// icp_target.cuh — device-resident search structure of one target cloud (the `Icp2d<'a>` /
// `Icp3d<'a>` value, src/lib.rs:91-102,133-144) and the stand-alone correspondence kernel
// (KdTree::search per query, src/lib.rs:121,164).
//
// Layout in HBM (SoA, sorted by cell, rows of cells contiguous along x):
//   x[m], y[m], (z[m])   f64   coordinates, bit-identical copies of the caller's points
//   idx[m]               u32   original index of the point at each sorted position
//   cstart[ncells + 1]   u32   first sorted position of every cell
// Algorithmic traffic of one search call: 8*DIM*Q (queries) + 4*Q (indices) + the structure
// read once = 20 B/query in 2-D, 28 B/query in 3-D for Q >> m.
#pragma once
#include "icp_fused.cuh"

namespace icp {

struct TargetBuild {
    const double* pts;     // AoS, caller's order
    uint32_t m;
    GridGeom g;
    uint32_t* cnt;         // [ncells] zeroed
    uint32_t* slot;        // [m]
    uint32_t* cstart;      // [ncells+1]
    double* x; double* y; double* z;
    uint32_t* idx;
    unsigned long long* bbox_keys;  // [6] min xyz, max xyz as order-preserving keys
};

template <int DIM>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, 1) target_bbox_kernel(TargetBuild b)
{
    const double inf = 1.0 / 0.0;
    double mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < b.m; j += gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double v = b.pts[(size_t)j * DIM + d];
            mn[d] = fmin(mn[d], v);
            mx[d] = fmax(mx[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        mn[d] = warp_min_f64(mn[d]);
        mx[d] = warp_max_f64(mx[d]);
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            atomicMin(&b.bbox_keys[d], (unsigned long long)f64_order_key(mn[d]));
            atomicMax(&b.bbox_keys[3 + d], (unsigned long long)f64_order_key(mx[d]));
        }
    }
}

template <int DIM>
ICP_DEV uint32_t target_cell_of(const TargetBuild& b, uint32_t j)
{
    const GridGeom& g = b.g;
    int cx = grid_coord(b.pts[(size_t)j * DIM + 0], g.ox, g.inv_h, g.gx);
    int cy = grid_coord(b.pts[(size_t)j * DIM + 1], g.oy, g.inv_h, g.gy);
    int cz = (DIM == 3) ? grid_coord(b.pts[(size_t)j * DIM + 2], g.oz, g.inv_h, g.gz) : 0;
    return (uint32_t)((cz * g.gy + cy) * g.gx + cx);
}

template <int DIM>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, 1) target_count_kernel(TargetBuild b)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < b.m; j += gridDim.x * blockDim.x)
        b.slot[j] = atomicAdd(&b.cnt[target_cell_of<DIM>(b, j)], 1u);
}

// single CTA: exclusive scan of cnt[ncells] into cstart[ncells+1]
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(1024, 1) target_scan_kernel(TargetBuild b, uint32_t ncells)
{
    __shared__ uint32_t warp_tot[32];
    const int lane = lane_id(), warp = warp_id();
    const uint32_t nw = blockDim.x / 32;
    const uint32_t chunk = (((ncells + nw - 1) / nw) + 31u) & ~31u;
    const uint32_t c0 = (uint32_t)warp * chunk;
    uint32_t s = 0;
    for (uint32_t c = c0 + lane; c < c0 + chunk; c += 32) s += (c < ncells) ? b.cnt[c] : 0u;
    s = warp_sum_u32(s);
    if (lane == 0) warp_tot[warp] = s;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    for (uint32_t c = c0 + lane; c < c0 + chunk; c += 32) {
        uint32_t v = (c < ncells) ? b.cnt[c] : 0u;
        uint32_t incl = warp_inclusive_scan_u32(v);
        if (c < ncells) b.cstart[c] = base + incl - v;
        base += __shfl_sync(ICP_FULL_MASK, incl, 31);
    }
    if (threadIdx.x == 0) b.cstart[ncells] = b.m;
}

template <int DIM>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, 1) target_scatter_kernel(TargetBuild b)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < b.m; j += gridDim.x * blockDim.x) {
        uint32_t pos = b.cstart[target_cell_of<DIM>(b, j)] + b.slot[j];
        b.x[pos] = b.pts[(size_t)j * DIM + 0];
        b.y[pos] = b.pts[(size_t)j * DIM + 1];
        if (DIM == 3) b.z[pos] = b.pts[(size_t)j * DIM + 2];
        b.idx[pos] = j;
    }
}

// One thread per query.  queries: AoS f64.  out_idx: original index of the exact nearest target
// point (lowest index on ties).
template <int DIM>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, 2) nn_search_kernel(TargetView tv, const double* queries,
                                                          unsigned long long nq, int32_t* out_idx)
{
    GridView<DIM, uint32_t, uint32_t> gv;
    gv.x = tv.x; gv.y = tv.y; gv.z = tv.z; gv.idx = tv.idx; gv.cstart = tv.cstart; gv.g = tv.g;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        double qx, qy, qz = 0.0;
        if (DIM == 2) {
            const double2 v = reinterpret_cast<const double2*>(queries)[q];
            qx = v.x; qy = v.y;
        } else {
            qx = queries[q * 3 + 0]; qy = queries[q * 3 + 1]; qz = queries[q * 3 + 2];
        }
        NnResult<DIM, uint32_t, uint32_t> r = grid_nn_search(gv, qx, qy, qz);
        out_idx[q] = r.idx;
    }
}

}  // namespace icp
