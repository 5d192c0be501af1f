// icp_target.cuh — device-resident search structure of one target cloud (the `Icp2d<'a>` /
// `Icp3d<'a>` value, src/lib.rs:91-102,133-144) and the stand-alone correspondence kernel
// (KdTree::search per query, src/lib.rs:121,164).
//
// Layout in HBM (SoA, sorted by cell, rows of cells contiguous along x):
//   x[m], y[m], (z[m])   f64   coordinates, bit-identical copies of the caller's points
//   idx[m]               u32   original index of the point at each sorted position
//   cstart[ncells + 1]   u32   first sorted position of every cell
//   2-D only (icp_target_nn.cuh):
//   pf[m]                f32x2 the point in cell units relative to the centre of its 32x32-cell block
//   seed[ncells]         u32   sorted position of a point near each cell (jump flooding)
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
    float2* pf;            // null: no binary32 copies (3-D, degenerate geometry)
    uint32_t* where;       // nullable [m]: sorted position of every point (incremental update: seeds are carried over)
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

// single CTA (128 B of dynamic shared memory): exclusive scan of cnt[ncells] into cstart[ncells+1]
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(1024, 1) target_scan_kernel(TargetBuild b, uint32_t ncells)
{
    ICP_SHARED_DYN(scan_smem);   // 32 x u32
    uint32_t* warp_tot = reinterpret_cast<uint32_t*>(scan_smem);
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
        if (b.where) b.where[j] = pos;
        if (DIM == 2 && b.pf) {
            const GridGeom& g = b.g;
            const double px = b.pts[(size_t)j * 2 + 0], py = b.pts[(size_t)j * 2 + 1];
            const int cx = grid_coord(px, g.ox, g.inv_h, g.gx);
            const int cy = grid_coord(py, g.oy, g.inv_h, g.gy);
            float2 v;
            v.x = (float)(nn_cell_coord(px, g.ox, g.inv_h) - nn_block_centre(cx));
            v.y = (float)(nn_cell_coord(py, g.oy, g.inv_h) - nn_block_centre(cy));
            b.pf[pos] = v;
        }
    }
}

// Seeds by jump flooding (2-D).  seed_init: a cell's own first point, or none.  seed_pass(step):
// every cell keeps the closest (to its centre, binary32 is plenty) of its own seed and the seeds of
// the 8 cells `step` away.  Ping-pong buffers keep the outcome independent of the schedule.  Seeds
// only bound the search disc (icp_target_nn.cuh): their quality affects speed, never results.
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, 1) target_seed_init_kernel(const uint32_t* cstart, uint32_t ncells, uint32_t* seed)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += gridDim.x * blockDim.x)
        seed[c] = (cstart[c + 1] > cstart[c]) ? cstart[c] : kNnNoSeed;
}

ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, 1) target_seed_pass_kernel(GridGeom g, const double* x, const double* y,
                                                                   const uint32_t* in, uint32_t* out, int step)
{
    const uint32_t ncells = (uint32_t)(g.gx * g.gy);
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += gridDim.x * blockDim.x) {
        const int cy = (int)(c / (uint32_t)g.gx), cx = (int)(c - (uint32_t)cy * (uint32_t)g.gx);
        const double mx = g.ox + ((double)cx + 0.5) * g.h, my = g.oy + ((double)cy + 0.5) * g.h;
        uint32_t best = kNnNoSeed;
        float bd = 3.0e38f;
        for (int k = 0; k < 9; ++k) {   // k == 4 is the cell itself
            const int nx = cx + (k % 3 - 1) * step, ny = cy + (k / 3 - 1) * step;
            if (nx < 0 || ny < 0 || nx >= g.gx || ny >= g.gy) continue;
            const uint32_t s = in[(uint32_t)ny * (uint32_t)g.gx + (uint32_t)nx];
            if (s == kNnNoSeed) continue;
            const float ex = (float)(x[s] - mx), ey = (float)(y[s] - my);
            const float d = ex * ex + ey * ey;
            if (d < bd || (d == bd && s < best)) { bd = d; best = s; }
        }
        out[c] = best;
    }
}

// Incremental update (icp_b200_target_update): every cell keeps the point its seed named — old sorted position
// -> old index -> new index (a removed point hands over to its successor) -> new sorted position.
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, 1) target_seed_remap_kernel(const uint32_t* seed_old, uint32_t ncells, const uint32_t* idx_old,
                                                                    uint32_t m_old, const uint32_t* remap, const uint32_t* where,
                                                                    uint32_t m_new, uint32_t* seed_new)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += gridDim.x * blockDim.x) {
        const uint32_t s = seed_old[c];
        uint32_t out = kNnNoSeed;
        if (s < m_old) {
            const uint32_t ni = remap[idx_old[s]];
            if (ni < m_new) out = where[ni];
        }
        seed_new[c] = out;
    }
}

// ... and a cell that received an added point is seeded with it (any point of a cell bounds its queries' discs).
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, 1) target_seed_added_kernel(GridGeom g, const double* pts, uint32_t first, uint32_t m,
                                                                    const uint32_t* where, uint32_t* seed)
{
    for (uint32_t j = first + blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const int cx = grid_coord(pts[(size_t)j * 2 + 0], g.ox, g.inv_h, g.gx);
        const int cy = grid_coord(pts[(size_t)j * 2 + 1], g.oy, g.inv_h, g.gy);
        seed[(uint32_t)cy * (uint32_t)g.gx + (uint32_t)cx] = where[j];   // several added points in one cell: any of them
    }
}

// One thread per query.  queries: AoS f64.  out_idx: original index of the exact nearest target
// point (lowest index on ties).
// Occupancy and query order, measured on the 16.8 M scan queries (6.35 G queries/s as built: 64 registers,
// four 256-thread blocks per SM, plain grid stride): 40 / 32 registers (6 / 8 blocks) 5.52 / 5.62 G; 3 / 2
// blocks (capped with unused shared memory, which also shrinks L1) 4.95-4.33 / 3.70 G; a block working
// through 1080-8640 consecutive queries before moving on 5.2-5.6 G against 5.34 G for the same loop with
// plain stride (the nested loop alone costs 16 %).  The kernel wants all of L1 and exactly this occupancy.
#ifndef ICP_NN_MINB
#define ICP_NN_MINB 2
#endif
template <int DIM>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(256, ICP_NN_MINB) nn_search_kernel(TargetView tv, const double* queries,
                                                          unsigned long long nq, int32_t* out_idx)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        double qx, qy, qz = 0.0;
        if (DIM == 2) {
            const double2 v = reinterpret_cast<const double2*>(queries)[q];
            qx = v.x; qy = v.y;
        } else {
            qx = queries[q * 3 + 0]; qy = queries[q * 3 + 1]; qz = queries[q * 3 + 2];
        }
        // the reference would panic comparing NaN distances; a non-finite query gets index -1
        if (!(fabs(qx) <= kF64Max) || !(fabs(qy) <= kF64Max) || !(fabs(qz) <= kF64Max)) { out_idx[q] = -1; continue; }
        const NnResult<DIM, uint32_t, uint32_t> r = target_nn_search<DIM>(tv, qx, qy, qz, -1, false);
        out_idx[q] = r.idx;
    }
}

}  // namespace icp

namespace icp {

// Warp-cooperative search kernel (2-D targets with seeds and binary32 copies): warp w takes queries
// [32 w', 32 w' + 32) of the array — consecutive queries, i.e. consecutive beams when the queries are
// scans.  target_nn_search2_warp (icp_target_nn.cuh).  256 threads, 512 B of shared memory per warp.
constexpr int kNnWarpThreads = 256;
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(kNnWarpThreads, 4) nn_search_warp_kernel(TargetView tv, const double* queries,
                                                                           unsigned long long nq, int32_t* out_idx)
{
    ICP_SHARED_DYN(nn_smem);
    NnWarpScratch* ws = reinterpret_cast<NnWarpScratch*>(nn_smem) + warp_id();
    const int lane = lane_id();
    const unsigned long long groups = (nq + 31ull) / 32ull;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * (kNnWarpThreads / 32);
    for (unsigned long long g = (unsigned long long)blockIdx.x * (kNnWarpThreads / 32) + (unsigned long long)warp_id(); g < groups;
         g += nwarps) {
        unsigned long long q = g * 32ull + (unsigned long long)lane;
        const bool real = q < nq;
        if (!real) q = nq - 1ull;       // padding lanes repeat the last query
        const double qx = queries[q * 2ull], qy = queries[q * 2ull + 1ull];
        const int pos = target_nn_search2_warp(tv, qx, qy, ws);
        if (real) out_idx[q] = (int32_t)tv.idx[pos];
    }
}
constexpr uint32_t kNnWarpShared = (uint32_t)(sizeof(NnWarpScratch) * (kNnWarpThreads / 32));

// Group search kernel (2-D targets with seeds and binary32 copies): G lanes per query
// (target_nn_search2_group); a warp takes 32 / G consecutive queries at a time.
constexpr int kNnGroupThreads = 256;
template <int G>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(kNnGroupThreads, 4) nn_search_group_kernel(TargetView tv, const double* queries,
                                                                             unsigned long long nq, int32_t* out_idx)
{
    constexpr unsigned long long QPW = 32 / G;
    const int lane = lane_id();
    const unsigned long long chunks = (nq + QPW - 1ull) / QPW;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * (kNnGroupThreads / 32);
    for (unsigned long long c = (unsigned long long)blockIdx.x * (kNnGroupThreads / 32) + (unsigned long long)warp_id(); c < chunks;
         c += nwarps) {
        const unsigned long long q = c * QPW + (unsigned long long)(lane / G);
        bool active = q < nq;
        double qx = 0., qy = 0.;
        if (active) {
            const double2 v = reinterpret_cast<const double2*>(queries)[q];
            qx = v.x; qy = v.y;
        }
        // the reference would panic comparing NaN distances; a non-finite query gets index -1
        const bool finite = (fabs(qx) <= kF64Max) && (fabs(qy) <= kF64Max);
        const int pos = target_nn_search2_group<G>(tv, qx, qy, active && finite);
        if (active && (lane & (G - 1)) == 0) out_idx[q] = finite ? (int32_t)tv.idx[pos] : -1;
    }
}

}  // namespace icp
