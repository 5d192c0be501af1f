// icp_stream.cuh — the ICP loop of one LARGE scan pair in one CTA, state streamed through L2.
//
// Same algorithm and helpers as icp_fused.cuh (Icp2d::estimate src/lib.rs:105-130, Icp3d::estimate
// src/lib.rs:148-173 and everything they call), for scans that do not fit the registers / shared
// memory of one SM: the bundled 3-D frames hold ~27k points (examples/scan3d.rs:104-131), config 4
// of BASELINE.json 16k.  Per CTA a workspace slab in global memory (L2-resident: ~2 MB per CTA)
// holds
//   - the pair's own sorted cell list (x, y, (z), idx, cstart), built by the CTA itself
//     (Icp3d::new, lib.rs:139-144) unless a shared device target is used;
//   - per source point the transformed point and its matched target point (sx, sy | dx, dy).
// Every Gauss-Newton evaluation streams that state: two exact selections (median, MAD; FAST map
// with zoom, icp_cta.cuh) and one accumulation pass, reduced as the fixed warp tree of
// cta_tree_reduce_solve (default), or — StreamParams::ordered — in the reference's own order: the
// row terms (lib.rs:242-255) go to the slab, tiles of kStreamTile points are staged in shared memory
// (double buffered) and the chain warp of icp_fused.cuh adds them sequentially, carrying its twelve
// accumulators from tile to tile.  That mode gives bit-identical transforms for scans of any size
// (it is also the re-run path of large pairs the knife-edge guard flags).
#pragma once
#ifdef ICP_STREAM_TIMING
#include <cstdio>
#endif
#include "icp_fused.cuh"

namespace icp {

#ifndef ICP_STREAM_NN_GROUP
#define ICP_STREAM_NN_GROUP 8
#endif
constexpr int kStreamNnGroup = ICP_STREAM_NN_GROUP;   // lanes per query of the correspondence search
constexpr int kStreamHeavyRows = 48;   // a ball covering more (slab, row) pairs than this is searched by a whole warp
constexpr int kStreamTile = 256;   // points per shared-memory tile of the ordered chain

struct StreamLayout {   // byte offsets inside one CTA's slab
    uint64_t s0, s1;                       // double2[n_cap] each
    uint64_t spos;                         // u32[n_cap]: sorted position of each point's previous match
    uint64_t row_r, row_c, row_b, row_rho; // ordered mode: row terms, padded to whole tiles
    uint64_t gx, gy, gz, gidx, gslot;      // target copy sorted by cell
    uint64_t grec;                         // the same as array-of-structs records (GridRec): the warp search's view
    uint64_t grecf;                        // 3-D: binary32 records in cell units (GridRecF): the filter's view
    uint64_t heavy;                        // u32[n_cap]: source points whose search is left to a whole warp
    uint64_t gcnt, gcstart;
    uint64_t total;
};

ICP_HD uint64_t align256(uint64_t v) { return (v + 255ull) & ~255ull; }

ICP_HD StreamLayout stream_layout(int dim, uint32_t n_cap, uint32_t m_cap, uint32_t c_cap, bool shared_target,
                                  bool ordered = false)
{
    StreamLayout L;
    uint64_t o = 0;
    L.s0 = o; o = align256(o + 16ull * n_cap);
    L.s1 = o; o = align256(o + 16ull * n_cap);
    L.spos = o; o = align256(o + 4ull * n_cap);
    const uint64_t n_pad = ordered ? ((uint64_t)n_cap + kStreamTile - 1) / kStreamTile * kStreamTile : 0ull;
    L.row_r = o; o = align256(o + 16ull * n_pad);
    L.row_c = o; o = align256(o + 16ull * n_pad);
    L.row_b = o; o = align256(o + 16ull * n_pad);
    L.row_rho = o; o = align256(o + 8ull * n_pad);
    if (shared_target) { m_cap = 0; c_cap = 0; }
    L.gx = o; o = align256(o + 8ull * m_cap);
    L.gy = o; o = align256(o + 8ull * m_cap);
    L.gz = o; if (dim == 3) o = align256(o + 8ull * m_cap);
    L.gidx = o; o = align256(o + 4ull * m_cap);
    L.gslot = o; o = align256(o + 4ull * m_cap);
    L.grec = o; o = align256(o + (uint64_t)sizeof(GridRec) * m_cap);
    L.grecf = o; if (dim == 3) o = align256(o + (uint64_t)sizeof(GridRecF) * m_cap);
    L.heavy = o; o = align256(o + 4ull * n_cap);
    L.gcnt = o; o = align256(o + 4ull * c_cap);
    L.gcstart = o; o = align256(o + 4ull * (c_cap + 1ull));
    L.total = o;
    return L;
}

struct StreamParams {
    FusedParams f;          // same batch description as the fused kernel (workspace/ws_stride = slabs)
    uint32_t n_cap, m_cap, c_cap;
    uint32_t cells_per_point;   // cell budget of a pair's own grid = cells_per_point * m, at most c_cap (0: 2)
    uint32_t ordered;           // reference-order sums (ICP_B200_FLAG_ORDERED_SUMS) instead of the warp tree
    uint32_t nn_group;          // 1: kStreamNnGroup lanes per query (grid_nn_search_seeded_group); 0: one thread per query
    uint32_t nn_f32;            // 1: 3-D per-pair searches rank candidates on binary32 records first (grid_nn_search_seeded_f32)
};

// Phase clock of CTA 0 (debug builds, -DICP_STREAM_TIMING): thread 0 sums clock64() differences.
#ifdef ICP_STREAM_TIMING
#define ICP_ST_DECL long long st_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long st_t = clock64(), st_pairs = 0, st_evals = 0;
#define ICP_ST_MARK(q) { const long long st_now = clock64(); st_acc[q] += st_now - st_t; st_t = st_now; }
#else
#define ICP_ST_DECL
#define ICP_ST_MARK(q)
#endif

// Shared-memory tile of the ordered chain: the row arrays of kStreamTile points at the 16-byte bank
// phases chain_run's 128-bit loads expect (icp_fused.cuh fused_layout: r +0, b +16, c alone).
struct StreamTileLayout { uint32_t r, b, c, rho, bytes; };
ICP_HD StreamTileLayout stream_tile_layout()
{
    StreamTileLayout t;
    uint32_t o = 0;
    t.r = o; o = align128(o + 16u * kStreamTile);
    t.b = o + 16u; o = align128(t.b + 16u * kStreamTile);
    t.c = o; o = align128(o + 16u * kStreamTile);
    t.rho = o; o = align128(o + 8u * kStreamTile);
    t.bytes = o;
    return t;
}

// residual keys recomputed from the streamed state (lib.rs:34-36)
struct KeysStreamed {
    const double2* s0;   // (sx, sy)
    const double2* s1;   // (dx, dy)
    Transform Ti;
    ICP_DEV_MEMBER double2 get(int /*kk*/, int i) const
    {
        const double2 s = s0[i], d = s1[i];
        double tx_, ty_;
        transform_apply(Ti, s.x, s.y, &tx_, &ty_);
        double2 r;
        r.x = tx_ - d.x;
        r.y = ty_ - d.y;
        return r;
    }
};

constexpr uint32_t kStreamSharedBase = (uint32_t)((((sizeof(SelShared) + 15u) & ~15u) + sizeof(FusedCtl) + 127u) & ~127u);
// + ordered mode: two tiles and the constant table of the chain (K0 at phase +32, K1 at +48)
ICP_HD uint32_t stream_smem_bytes(bool ordered)
{
    if (!ordered) return kStreamSharedBase;
    return kStreamSharedBase + 2u * stream_tile_layout().bytes + align128(32u + 16u + 2u * 16u * (uint32_t)kChainUnroll);
}
constexpr uint32_t kStreamShared = kStreamSharedBase;

// Reference-order sums of one Gauss-Newton evaluation (lib.rs:242-255, 45-50) over the row arrays in
// the slab: tile t+1 is copied into shared memory while the chain warp (and the cost warp) consume
// tile t.  On return ctl->sums[0..11] = jtj, jtr and ctl->sums[12] = cost, visible to every thread.
// cost_only: only sums[12] (the final huber_error).  npad: a multiple of kStreamTile.
template <int THREADS>
ICP_DEV void stream_ordered_sums(unsigned char* tiles, const double2* ktab, const double2* R2, const double2* C2,
                                 const double2* B2, const double* RHO, int npad, bool cost_only, FusedCtl* ctl)
{
    constexpr int NW = THREADS / 32;
    const StreamTileLayout TL = stream_tile_layout();
    const int tid = (int)threadIdx.x, lane = lane_id(), warp = warp_id();
    const int ntiles = npad / kStreamTile;
    auto copy_tile = [&](int t) {
        unsigned char* buf = tiles + (size_t)(t & 1) * TL.bytes;
        const size_t p0 = (size_t)t * kStreamTile;
        for (int q = tid; q < kStreamTile; q += THREADS) {
            if (!cost_only) {
                reinterpret_cast<double2*>(buf + TL.r)[q] = R2[p0 + q];
                reinterpret_cast<double2*>(buf + TL.b)[q] = B2[p0 + q];
                reinterpret_cast<double2*>(buf + TL.c)[q] = C2[p0 + q];
            }
            reinterpret_cast<double*>(buf + TL.rho)[q] = RHO[p0 + q];
        }
    };
    double acc = 0.0, cost = 0.0;
    if (ntiles > 0) copy_tile(0);
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) copy_tile(t + 1);
        const unsigned char* buf = tiles + (size_t)(t & 1) * TL.bytes;
        if (warp == 0 && !cost_only)
            acc = chain_run(chain_role(lane, reinterpret_cast<const double2*>(buf + TL.r), reinterpret_cast<const double2*>(buf + TL.c),
                                       reinterpret_cast<const double2*>(buf + TL.b), ktab),
                            kStreamTile / kChainStage, acc);
        if (warp == (NW > 1 ? 1 : 0))
            cost = cost_run(smem_ref_of(reinterpret_cast<const double2*>(buf + TL.rho)), kStreamTile / 2, cost);
        __syncthreads();
    }
    if (warp == 0 && lane < 12 && !cost_only) ctl->sums[lane] = acc;
    if (warp == (NW > 1 ? 1 : 0) && lane == 0) ctl->sums[12] = cost;
    __syncthreads();
}

// CTA-wide exclusive scan of cnt[0..ncells) into cstart[0..ncells], one contiguous chunk per warp
template <int THREADS>
ICP_DEV void cta_scan_cells(const uint32_t* cnt, uint32_t* cstart, uint32_t ncells, uint32_t total, FusedCtl* ctl)
{
    constexpr uint32_t NW = THREADS / 32;
    const int lane = lane_id(), warp = warp_id();
    const uint32_t chunk = (((ncells + NW - 1) / NW) + 31u) & ~31u;
    const uint32_t c0 = (uint32_t)warp * chunk;
    uint32_t s = 0;
    for (uint32_t c = c0 + lane; c < c0 + chunk; c += 32) s += (c < ncells) ? cnt[c] : 0u;
    s = warp_sum_u32(s);
    if (lane == 0) ctl->warp_tot[warp] = s;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += ctl->warp_tot[w];
    for (uint32_t c = c0 + lane; c < c0 + chunk; c += 32) {
        const uint32_t v = (c < ncells) ? cnt[c] : 0u;
        const uint32_t incl = warp_inclusive_scan_u32(v);
        if (c < ncells) cstart[c] = base + incl - v;
        base += __shfl_sync(ICP_FULL_MASK, incl, 31);
    }
    if (threadIdx.x == 0) cstart[ncells] = total;
    __syncthreads();
}

// two CTAs per SM (64 registers per thread, some spills in the selection): the search is bound by the
// latency of its dependent L2 loads, so the second CTA's 16 warps are worth more than the registers —
// 3.88k against 2.98k pairs/s on config 4 (profiles/r02_config4.md)
#ifndef ICP_STREAM_MINB
#define ICP_STREAM_MINB 2
#endif
template <int THREADS, int DIM, bool SHARED_TARGET>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(THREADS, ICP_STREAM_MINB) icp_stream_kernel(StreamParams sp)
{
    constexpr int NW = THREADS / 32;
    static_assert(NW <= kMaxWarps, "FusedCtl holds per-warp slots for kMaxWarps warps");
    const FusedParams& p = sp.f;
    ICP_SHARED_DYN(smem);
    SelShared* sel = reinterpret_cast<SelShared*>(smem);
    FusedCtl* ctl = reinterpret_cast<FusedCtl*>(smem + ((sizeof(SelShared) + 15u) & ~15u));
    const int tid = (int)threadIdx.x, lane = lane_id(), warp = warp_id();
    const bool ordered = sp.ordered != 0;
    const StreamLayout L = stream_layout(DIM, sp.n_cap, sp.m_cap, sp.c_cap, SHARED_TARGET, ordered);
    unsigned char* slab = p.workspace + (size_t)blockIdx.x * p.ws_stride;
    double2* S0 = reinterpret_cast<double2*>(slab + L.s0);
    double2* S1 = reinterpret_cast<double2*>(slab + L.s1);
    uint32_t* spos = reinterpret_cast<uint32_t*>(slab + L.spos);
    double2* R2 = reinterpret_cast<double2*>(slab + L.row_r);
    double2* C2 = reinterpret_cast<double2*>(slab + L.row_c);
    double2* B2 = reinterpret_cast<double2*>(slab + L.row_b);
    double* RHO = reinterpret_cast<double*>(slab + L.row_rho);
    unsigned char* tiles = smem + kStreamSharedBase;
    double2* ktab = reinterpret_cast<double2*>(tiles + 2u * stream_tile_layout().bytes + 32u);

    for (int q = tid; q < 2 * kSelHistPad; q += THREADS) (&sel->hist[0][0])[q] = 0u;
    int hist_par = 0;
    fused_launch_entry(p);
    __syncthreads();
    ICP_ST_DECL

    for (;;) {
        if (tid == 0) { ctl->pair = fused_next_pair(p); ctl->knife = 0u; }
        __syncthreads();
        const uint32_t pair = ctl->pair;
        if (pair >= p.pair_end) break;
#ifdef ICP_STREAM_TIMING
        ++st_pairs;
#endif
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
        if ((uint32_t)n > sp.n_cap || n > 65535 || (!SHARED_TARGET && (uint32_t)m > sp.m_cap)) status = kPairTooLarge;
        {
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
                    for (uint32_t it = 0; it < p.max_iter; ++it) p.out_inner[(size_t)pair * p.max_iter + it] = 0u;
            }
            __syncthreads();
            continue;
        }

        // ---- IcpD::new: sorted cell list of this pair's target in the slab
        GridView<DIM, uint32_t, uint32_t> gv;
        GridRec* grec = reinterpret_cast<GridRec*>(slab + L.grec);
        GridRecF* grecf = (DIM == 3 && !SHARED_TARGET && sp.nn_f32) ? reinterpret_cast<GridRecF*>(slab + L.grecf) : nullptr;
        if (SHARED_TARGET) {
            gv.x = p.target.x; gv.y = p.target.y; gv.z = p.target.z;
            gv.idx = p.target.idx; gv.cstart = p.target.cstart; gv.g = p.target.g;
        } else {
            double* gx = reinterpret_cast<double*>(slab + L.gx);
            double* gy = reinterpret_cast<double*>(slab + L.gy);
            double* gz = reinterpret_cast<double*>(slab + L.gz);
            uint32_t* gidx = reinterpret_cast<uint32_t*>(slab + L.gidx);
            uint32_t* gslot = reinterpret_cast<uint32_t*>(slab + L.gslot);
            uint32_t* gcnt = reinterpret_cast<uint32_t*>(slab + L.gcnt);
            uint32_t* gcstart = reinterpret_cast<uint32_t*>(slab + L.gcstart);
            gv.x = gx; gv.y = gy; gv.z = gz; gv.idx = gidx; gv.cstart = gcstart;
            gv.g = grid_make_geom(0, 0, 0, 0, 0, 0, DIM, 1);
            if (m > 0) {
                const double inf = 1.0 / 0.0;
                double mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
                for (int j = tid; j < m; j += THREADS) {
#pragma unroll
                    for (int d = 0; d < DIM; ++d) {
                        const double v = dst[(size_t)j * DIM + d];
                        mn[d] = fmin(mn[d], v);
                        mx[d] = fmax(mx[d], v);
                    }
                }
#pragma unroll
                for (int d = 0; d < DIM; ++d) { mn[d] = warp_min_f64(mn[d]); mx[d] = warp_max_f64(mx[d]); }
                if (lane == 0) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { ctl->red[warp][d] = mn[d]; ctl->red[warp][3 + d] = mx[d]; }
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
                long long budget = (long long)(sp.cells_per_point ? sp.cells_per_point : 2u) * m;
                if (budget < 64) budget = 64;
                if (budget > (long long)sp.c_cap) budget = (long long)sp.c_cap;
                const GridGeom g = grid_make_geom(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], DIM, (int)budget);
                gv.g = g;
                const uint32_t ncells = (uint32_t)(g.gx * g.gy * g.gz);
                for (uint32_t c = (uint32_t)tid; c < ncells; c += THREADS) gcnt[c] = 0u;
                __syncthreads();
                auto cell_of = [&](int j) {
                    const int cx = grid_coord(dst[(size_t)j * DIM + 0], g.ox, g.inv_h, g.gx);
                    const int cy = grid_coord(dst[(size_t)j * DIM + 1], g.oy, g.inv_h, g.gy);
                    const int cz = (DIM == 3) ? grid_coord(dst[(size_t)j * DIM + 2], g.oz, g.inv_h, g.gz) : 0;
                    return (uint32_t)((cz * g.gy + cy) * g.gx + cx);
                };
                for (int j = tid; j < m; j += THREADS) gslot[j] = atomicAdd(&gcnt[cell_of(j)], 1u);
                __syncthreads();
                cta_scan_cells<THREADS>(gcnt, gcstart, ncells, (uint32_t)m, ctl);
                for (int j = tid; j < m; j += THREADS) {
                    const uint32_t pos = gcstart[cell_of(j)] + gslot[j];
                    gx[pos] = dst[(size_t)j * DIM + 0];
                    gy[pos] = dst[(size_t)j * DIM + 1];
                    if (DIM == 3) gz[pos] = dst[(size_t)j * DIM + 2];
                    gidx[pos] = (uint32_t)j;
                    GridRec rc;
                    rc.x = dst[(size_t)j * DIM + 0]; rc.y = dst[(size_t)j * DIM + 1];
                    rc.z = (DIM == 3) ? dst[(size_t)j * DIM + 2] : 0.0;
                    rc.idx = (uint32_t)j; rc.pad = 0u;
                    grec[pos] = rc;
                    if (DIM == 3 && grecf) {   // the same coordinates the walk gives a query (nn_cell_coord), rounded once
                        GridRecF rf;
                        rf.x = (float)nn_cell_coord(rc.x, g.ox, g.inv_h);
                        rf.y = (float)nn_cell_coord(rc.y, g.oy, g.inv_h);
                        rf.z = (float)nn_cell_coord(rc.z, g.oz, g.inv_h);
                        rf.idx = (uint32_t)j;
                        grecf[pos] = rf;
                    }
                }
                __syncthreads();
            }
        }

        ICP_ST_MARK(0)
        double est_med[2] = {0., 0.}, est_sigma[2] = {0., 0.};
        uint64_t H = 0;  // meaningful in thread 0
        const int npad = ((n + kStreamTile - 1) / kStreamTile) * kStreamTile;   // ordered mode: whole tiles
        const uint64_t corr_base = (p.shared_src ? (uint64_t)pair * (uint64_t)n : (uint64_t)p.src_off[pair]) *
                                   (p.corr_mode == 2 ? p.max_iter : 1u);

        for (uint32_t it = 0; it < p.max_iter; ++it) {
            // ---- transform (lib.rs:113-116 / transform_xy lib.rs:52-57) + correspondence search
            uint64_t hpart = 0;
            ICP_ST_MARK(6)
            if (SHARED_TARGET || !sp.nn_group) {
                // one thread per point; a point whose seeded ball is large (far previous match: a source
                // point without a counterpart) is put on a list and searched afterwards by a whole warp,
                // rows spread over the lanes, instead of stalling the other 31 lanes of this warp
                uint32_t* heavy = reinterpret_cast<uint32_t*>(slab + L.heavy);
                if (tid == 0) ctl->heavy_n = 0u;
                __syncthreads();
                auto finish = [&](int i, double2 s, const NnResult<DIM, uint32_t, uint32_t>& r) {
                    spos[i] = (uint32_t)r.pos;
                    double2 d;
                    if (SHARED_TARGET) { d.x = gv.x[r.pos]; d.y = gv.y[r.pos]; }   // get_xy (lib.rs:86-89)
                    else { const GridRec m = grec[r.pos]; d.x = m.x; d.y = m.y; }
                    S0[i] = s;
                    S1[i] = d;
                    hpart += mix64(((uint64_t)(uint32_t)i << 32) | (uint64_t)(uint32_t)r.idx);
                    if (p.out_corr) {
                        if (p.corr_mode == 2) p.out_corr[corr_base + (uint64_t)it * (uint64_t)n + (uint64_t)i] = r.idx;
                        else if (p.corr_mode == 1 && (it + 1 == p.max_iter || p.fixed_point)) p.out_corr[corr_base + (uint64_t)i] = r.idx;
                    }
                };
                // (visiting the points in the order of their own cell's occupancy, to even out the lanes of a
                // warp, was measured and dropped: 2.69k against 2.98k pairs/s on config 4 — consecutive
                // points share their candidates' cache lines, and that coherence is worth more)
                for (int i = tid; i < n; i += THREADS) {
                    double2 s;
                    transform_apply(T, src[(size_t)i * DIM + 0], src[(size_t)i * DIM + 1], &s.x, &s.y);
                    const double pz = (DIM == 3) ? src[(size_t)i * DIM + 2] : 0.0;
                    // from the second outer iteration on the previous match seeds the search (icp_target_nn.cuh)
                    const int seed = (it > 0) ? (int)spos[i] : -1;
                    if (!SHARED_TARGET && seed >= 0 && grid_ball_rows(gv, grec, s.x, s.y, pz, seed) > kStreamHeavyRows) {
                        S0[i] = s;
                        heavy[atomicAdd(&ctl->heavy_n, 1u)] = (uint32_t)i;
                        continue;
                    }
                    const NnResult<DIM, uint32_t, uint32_t> r =
                        SHARED_TARGET ? target_nn_search<DIM>(p.target, s.x, s.y, pz, seed)
                                      : grid_nn_search_auto(gv, s.x, s.y, pz, seed, grec, grecf);
                    finish(i, s, r);
                }
                __syncthreads();
                if (!SHARED_TARGET) {
                    const uint32_t nh = ctl->heavy_n;
                    for (uint32_t k = (uint32_t)warp; k < nh; k += NW) {
                        const int i = (int)heavy[k];
                        const double2 s = S0[i];
                        const double pz = (DIM == 3) ? src[(size_t)i * DIM + 2] : 0.0;
                        const int seed = (int)spos[i];
                        __syncwarp();
                        bool ok = true;
                        NnResult<DIM, uint32_t, uint32_t> r = grid_nn_search_seeded_group<DIM, 32>(gv, grec, s.x, s.y, pz, seed, true, &ok);
                        if (lane == 0) {
                            if (!ok) r = grid_nn_search(gv, s.x, s.y, pz);
                            finish(i, s, r);
                        }
                        __syncwarp();
                    }
                }
            } else {
                // pass 1, one thread per point: the transformed point and the seed of its search — the
                // previous match, in the first iteration the first point of the query's own cell
                for (int i = tid; i < n; i += THREADS) {
                    double2 s;
                    transform_apply(T, src[(size_t)i * DIM + 0], src[(size_t)i * DIM + 1], &s.x, &s.y);
                    S0[i] = s;
                    if (it == 0) {
                        const GridGeom& g = gv.g;
                        const int cx = grid_coord(s.x, g.ox, g.inv_h, g.gx);
                        const int cy = grid_coord(s.y, g.oy, g.inv_h, g.gy);
                        const int cz = (DIM == 3) ? grid_coord(src[(size_t)i * DIM + 2], g.oz, g.inv_h, g.gz) : 0;
                        const int c = (cz * g.gy + cy) * g.gx + cx;
                        const uint32_t c0 = gv.cstart[c], c1 = gv.cstart[c + 1];
                        spos[i] = c1 > c0 ? c0 : 0xffffffffu;
                    }
                }
                __syncthreads();
                ICP_ST_MARK(1)
                // pass 2, kNnGroup lanes per point (32 / kNnGroup points per warp at a time): exact ball
                // walk around the seed (grid_nn_search_seeded_group); a point without a seed (empty own
                // cell, first iteration) or outside the walk's range takes the ring search
                constexpr int G = kStreamNnGroup, QPW = 32 / G;
                const int gq = lane / G;
                for (int i0 = warp * QPW; i0 < n; i0 += NW * QPW) {
                    const int i = i0 + gq;
                    const bool have = i < n;
                    double2 s;
                    s.x = 0.; s.y = 0.;
                    double pz = 0.0;
                    uint32_t seed = 0xffffffffu;
                    if (have) {
                        s = S0[i];
                        pz = (DIM == 3) ? src[(size_t)i * DIM + 2] : 0.0;   // transform_xy keeps z (lib.rs:52-57)
                        seed = spos[i];
                    }
                    __syncwarp();   // every lane has read its seed before a group leader replaces it
                    bool ok = true;
                    NnResult<DIM, uint32_t, uint32_t> r =
                        grid_nn_search_seeded_group<DIM, G>(gv, grec, s.x, s.y, pz, (int)seed, have && seed != 0xffffffffu, &ok);
                    if (have && (lane % G) == 0) {
                        if (seed == 0xffffffffu || !ok) r = grid_nn_search(gv, s.x, s.y, pz);
                        spos[i] = (uint32_t)r.pos;
                        const GridRec m = grec[r.pos];
                        double2 d;
                        d.x = m.x; d.y = m.y;   // get_xy (lib.rs:86-89)
                        S1[i] = d;
                        hpart += mix64(((uint64_t)(uint32_t)i << 32) | (uint64_t)(uint32_t)r.idx);
                        if (p.out_corr) {
                            if (p.corr_mode == 2) p.out_corr[corr_base + (uint64_t)it * (uint64_t)n + (uint64_t)i] = r.idx;
                            else if (p.corr_mode == 1 && (it + 1 == p.max_iter || p.fixed_point)) p.out_corr[corr_base + (uint64_t)i] = r.idx;
                        }
                    }
                    __syncwarp();
                }
            }
            hpart = warp_sum_u64(hpart);
            if (lane == 0) ctl->hash_part[warp] = hpart;
            __syncthreads();   // state of every point is visible to the whole CTA
            ICP_ST_MARK(2)
            uint64_t hterm = 0;
            if (tid == 0) {
                for (int w = 0; w < NW; ++w) hterm += ctl->hash_part[w];
                H = H * 0x100000001B3ull + hterm;
            }

            // ---- estimate_transform (lib.rs:59-84)
            Transform Ti = transform_identity();
            double prev_error = kF64Max;
            uint32_t accepted = 0, why = kBreakCap;
            if (n < 2) {
                why = kBreakNone;
            } else {
                for (int inner = 0; inner < kInnerMaxIter; ++inner) {
                    KeysStreamed ks;
                    ks.s0 = S0; ks.s1 = S1; ks.Ti = Ti;
                    double sigma[2];
                    hist_par = cta_calc_stddevs<THREADS, 0>(sel, ks, n, est_med, est_sigma, hist_par, sigma);
                    ICP_ST_MARK(3)
#ifdef ICP_STREAM_TIMING
                    ++st_evals;
#endif
                    const double g0 = 1. / sigma[0], g1 = 1. / sigma[1];
                    if (ordered) {
                        // row weights (lib.rs:245-250), Jacobian third column (lib.rs:176-184), Huber cost
                        // terms (lib.rs:47-48) -> slab; zero rows pad the last tile (they add +0.0)
                        for (int i = tid; i < npad; i += THREADS) {
                            double2 rr, cc, bb;
                            double rh = 0.;
                            rr.x = 0.; rr.y = 0.; cc.x = 0.; cc.y = 0.; bb.x = 0.; bb.y = 0.;
                            if (i < n) {
                                const double2 s = S0[i];
                                rr = ks.get(0, i);
                                if (sigma[0] != 0.) cc.x = huber_drho_of_residual(rr.x) * g0;
                                if (sigma[1] != 0.) cc.y = huber_drho_of_residual(rr.y) * g1;
                                const double a0 = -s.y, a1 = s.x;
                                bb.x = Ti.r00 * a0 + Ti.r01 * a1;
                                bb.y = Ti.r10 * a0 + Ti.r11 * a1;
                                rh = huber_rho(rr.x * rr.x + rr.y * rr.y);
                            }
                            R2[i] = rr; C2[i] = cc; B2[i] = bb; RHO[i] = rh;
                        }
                        if (tid < 2 * kChainUnroll) {
                            double2 kv;
                            if (tid < kChainUnroll) { kv.x = Ti.r00; kv.y = Ti.r10; }
                            else { kv.x = Ti.r01; kv.y = Ti.r11; }
                            ktab[tid < kChainUnroll ? tid : tid - kChainUnroll + kChainK1] = kv;
                        }
                        __syncthreads();
                        stream_ordered_sums<THREADS>(tiles, ktab, R2, C2, B2, RHO, npad, false, ctl);
                        double jtj[9], jtr[3], delta[3];
#pragma unroll
                        for (int q = 0; q < 9; ++q) jtj[q] = ctl->sums[q];
#pragma unroll
                        for (int q = 0; q < 3; ++q) jtr[q] = ctl->sums[9 + q];
                        const double error = ctl->sums[12];
                        __syncthreads();   // sums are rewritten by the next evaluation
                        // solve + break tests + update: every thread, identically (lib.rs:67-81)
                        if (!solve_normal_equations(jtj, jtr, delta)) { why = kBreakNone; break; }
                        if ((delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2] < kDeltaNormThreshold) {
                            why = kBreakDelta;
                            break;
                        }
                        if (error > prev_error) { why = kBreakError; break; }
                        prev_error = error;
                        Ti = transform_mul(transform_new(delta[0], delta[1], delta[2]), Ti);
                        ++accepted;
                        continue;
                    }
                    double acc[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) acc[q] = 0.;
                    for (int i = tid; i < n; i += THREADS) {
                        const double2 s = S0[i];
                        tree_accumulate_point(acc, Ti, s.x, s.y, ks.get(0, i), sigma, g0, g1);
                    }
                    cta_tree_reduce_solve<NW>(ctl, acc, Ti, prev_error, p.guard);
                    ICP_ST_MARK(4)
                    const uint32_t decision = (uint32_t)ctl->bcast[7];
                    if (decision != kBreakCap) { why = decision; break; }
                    prev_error = ctl->bcast[6];
                    Ti.r00 = ctl->bcast[0]; Ti.r01 = ctl->bcast[1]; Ti.r10 = ctl->bcast[2];
                    Ti.r11 = ctl->bcast[3]; Ti.tx = ctl->bcast[4];  Ti.ty = ctl->bcast[5];
                    ++accepted;
                }
            }
            const bool repeats = fused_repeats(p, accepted, it);   // icp_fused.cuh: the remaining iterations are copies
            if (p.out_cost && (it + 1 == p.max_iter || repeats)) {
                // huber_error (lib.rs:45-50) of the inner solve's final transform over this (the last)
                // iteration's transformed source / matched target points
                __syncthreads();
                KeysStreamed kf;
                kf.s0 = S0; kf.s1 = S1; kf.Ti = Ti;
                if (ordered) {
                    for (int i = tid; i < npad; i += THREADS) {
                        double rh = 0.;
                        if (i < n) {
                            const double2 rr = kf.get(0, i);
                            rh = huber_rho(rr.x * rr.x + rr.y * rr.y);
                        }
                        RHO[i] = rh;
                    }
                    __syncthreads();
                    stream_ordered_sums<THREADS>(tiles, ktab, R2, C2, B2, RHO, npad, true, ctl);
                    if (tid == 0) p.out_cost[pair] = npad > 0 ? ctl->sums[12] : 0.;
                } else {
                    double e = 0.;
                    for (int i = tid; i < n; i += THREADS) {
                        const double2 rr = kf.get(0, i);
                        e = e + huber_rho(rr.x * rr.x + rr.y * rr.y);
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
            __syncthreads();
            if (repeats) break;
        }
        if (tid == 0) {
            p.out_T[(size_t)pair * 6 + 0] = T.r00; p.out_T[(size_t)pair * 6 + 1] = T.r01;
            p.out_T[(size_t)pair * 6 + 2] = T.r10; p.out_T[(size_t)pair * 6 + 3] = T.r11;
            p.out_T[(size_t)pair * 6 + 4] = T.tx;  p.out_T[(size_t)pair * 6 + 5] = T.ty;
            if (p.out_status) p.out_status[pair] = kPairOk;
            if (p.out_hash) p.out_hash[pair] = H;
            if (p.out_cost && p.max_iter == 0) p.out_cost[pair] = 0.;
            if (!ordered) fused_report_knife(p, ctl, pair);
        }
        __syncthreads();
    }
    if (tid == 0) fused_launch_exit(p);
#ifdef ICP_STREAM_TIMING
    if (blockIdx.x == 0 && tid == 0 && st_pairs > 0)
        printf("stream timing CTA0: pairs %lld evals %lld | cycles/pair: setup+build %lld seeds %lld search %lld selection %lld accumulate+solve %lld other %lld\n",
               st_pairs, st_evals, st_acc[0] / st_pairs, st_acc[1] / st_pairs, st_acc[2] / st_pairs, st_acc[3] / st_pairs,
               st_acc[4] / st_pairs, st_acc[6] / st_pairs);
#endif
}

}  // namespace icp
