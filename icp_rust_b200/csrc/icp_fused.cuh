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
// Shared memory is one region used as {sorted target, cell table} during search and as
// {r, c, b, rho} row arrays during IRLS.
#pragma once
#include "icp_grid.cuh"

namespace icp {

constexpr int kChainUnroll = 8;  // rows arrays are zero-padded to a multiple of this many points
constexpr int kMaxWarps = 32;

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
    double red[kMaxWarps][6];
    uint64_t hash_part[kMaxWarps];
    uint32_t warp_tot[kMaxWarps];
    uint32_t pair;
};

ICP_HD uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

// Shared target (Icp2d::new analogue kept on the device, reused by many estimates)
struct TargetView {
    const double* x;
    const double* y;
    const double* z;
    const uint32_t* idx;
    const uint32_t* cstart;
    GridGeom g;
    uint32_t m;
};

ICP_HD FusedLayout fused_layout(int dim, int ncap, int mcap, int cell_budget, bool shared_target)
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
    L.bt_cnt = o; o = align16(o + 4u * (uint32_t)cell_budget);
    L.bt_cell = o; o = align16(o + 2u * (uint32_t)mcap);
    L.bt_slot = o; o = align16(o + 2u * (uint32_t)mcap);
    uint32_t nn_total = o;
    o = 0;
    L.gn_r = o; o = align16(o + 16u * (uint32_t)ncap);
    L.gn_c = o; o = align16(o + 16u * (uint32_t)ncap);
    L.gn_b = o; o = align16(o + 16u * (uint32_t)ncap);
    L.gn_rho = o; o = align16(o + 8u * (uint32_t)ncap);
    L.gn_const = o; o = align16(o + 8u * 6u * 2u * (uint32_t)kChainUnroll);
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
};

// ---------------------------------------------------------------------------------------------
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
// Sequential accumulation (lib.rs:242-255 and lib.rs:45-50).  Lane L of the chain warp owns one
// accumulator and adds its terms in the reference's order.  Term of row (i, j):
//     (X * Y) * Z   with  X = c_ij = w_ij * g_j,   Y = J_ij[a],   Z = J_ij[b] or r_ij
// J_ij = [R_j0, R_j1, b_ij].  Point-dependent factors are read from the row arrays (stride two
// doubles per point); uniform factors come from a table whose entries repeat, so that every
// lane executes the same loads with the same immediates.
struct ChainPtrs {
    const double *x0, *y0, *z0, *x1, *y1, *z1;
    int ix0, iy0, iz0, ix1, iy1, iz1;   // pointer advance per block of kChainUnroll points
};

ICP_DEV double chain_run(ChainPtrs p, int nblocks)
{
    double acc = 0.0;
    for (int blk = 0; blk < nblocks; ++blk) {
#pragma unroll
        for (int u = 0; u < kChainUnroll; ++u) {
            acc = acc + (p.x0[2 * u] * p.y0[2 * u]) * p.z0[2 * u];
            acc = acc + (p.x1[2 * u] * p.y1[2 * u]) * p.z1[2 * u];
        }
        p.x0 += p.ix0; p.y0 += p.iy0; p.z0 += p.iz0;
        p.x1 += p.ix1; p.y1 += p.iy1; p.z1 += p.iz1;
    }
    return acc;
}

ICP_DEV ChainPtrs chain_role(int lane, const double* r2, const double* c2, const double* b2,
                             const double* ktab)
{
    // ktab[k][2*kChainUnroll], k: 0..3 = R00 R01 R10 R11, 4 = 1.0, 5 = 0.0
    constexpr int KS = 2 * kChainUnroll;
    constexpr int VAR = 2 * kChainUnroll;
    ChainPtrs p;
    auto jac = [&](int j, int col, const double** ptr, int* inc) {
        if (col < 2) { *ptr = ktab + (2 * j + col) * KS; *inc = 0; }
        else { *ptr = b2 + j; *inc = VAR; }
    };
    if (lane < 12) {
        const int a = (lane < 9) ? lane / 3 : lane - 9;
        p.x0 = c2; p.x1 = c2 + 1; p.ix0 = VAR; p.ix1 = VAR;
        jac(0, a, &p.y0, &p.iy0);
        jac(1, a, &p.y1, &p.iy1);
        if (lane < 9) {
            const int b = lane % 3;
            jac(0, b, &p.z0, &p.iz0);
            jac(1, b, &p.z1, &p.iz1);
        } else {
            p.z0 = r2; p.z1 = r2 + 1; p.iz0 = VAR; p.iz1 = VAR;
        }
    } else {
        p.x0 = p.x1 = ktab + 5 * KS;
        p.y0 = p.y1 = p.z0 = p.z1 = ktab + 4 * KS;
        p.ix0 = p.ix1 = p.iy0 = p.iy1 = p.iz0 = p.iz1 = 0;
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
template <int THREADS, int PPT, int DIM, bool SHARED_TARGET, int MINB>
ICP_GLOBAL void ICP_LAUNCH_BOUNDS(THREADS, MINB) icp_fused_kernel(FusedParams p)
{
    constexpr int NW = THREADS / 32;
    constexpr int NCAP = THREADS * PPT;
    static_assert(NCAP % kChainUnroll == 0, "row arrays are padded to the chain unroll");
    ICP_SHARED_DYN(smem);
    const FusedLayout L = fused_layout(DIM, NCAP, p.mcap, p.cell_budget, SHARED_TARGET);
    unsigned char* U = smem;
    SelShared* sel = reinterpret_cast<SelShared*>(smem + L.off_sel);
    FusedCtl* ctl = reinterpret_cast<FusedCtl*>(smem + L.off_ctl);
    double* r2 = reinterpret_cast<double*>(U + L.gn_r);
    double* c2 = reinterpret_cast<double*>(U + L.gn_c);
    double* b2 = reinterpret_cast<double*>(U + L.gn_b);
    double* rho = reinterpret_cast<double*>(U + L.gn_rho);
    double* ktab = reinterpret_cast<double*>(U + L.gn_const);
    double2* r2v = reinterpret_cast<double2*>(r2);
    double2* c2v = reinterpret_cast<double2*>(c2);
    double2* b2v = reinterpret_cast<double2*>(b2);

    const int tid = (int)threadIdx.x, lane = lane_id(), warp = warp_id();
    unsigned char* slab = p.workspace + (size_t)blockIdx.x * p.ws_stride;

    for (int q = tid; q < 2 * 2 * kSelHistPad; q += THREADS) (&sel->hist[0][0][0])[q] = 0u;
    int hist_par = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) ctl->pair = p.pair_begin + atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const uint32_t pair = ctl->pair;
        if (pair >= p.pair_end) break;
        const int n = (int)(p.src_off[pair + 1] - p.src_off[pair]);
        const double* src = p.src + (size_t)p.src_off[pair] * DIM;
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
                for (uint32_t o = (uint32_t)tid * 16u; o < L.nn_bytes; o += THREADS * 16u)
                    *reinterpret_cast<uint4*>(slab + o) = *reinterpret_cast<const uint4*>(U + o);
            }
        } else {
            geom = p.target.g;
        }

        double est_med[2] = {0., 0.}, est_sigma[2] = {0., 0.};
        uint64_t H = 0;  // meaningful in thread 0
        const int npad = ((n + kChainUnroll - 1) / kChainUnroll) * kChainUnroll;
        const uint64_t corr_base = (uint64_t)p.src_off[pair] * (p.corr_mode == 2 ? p.max_iter : 1u);

        for (uint32_t it = 0; it < p.max_iter; ++it) {
            // ---- stage the sorted target (the union region was overwritten by the IRLS rows)
            if (!SHARED_TARGET && it > 0 && m > 0) {
                for (uint32_t o = (uint32_t)tid * 16u; o < L.nn_bytes; o += THREADS * 16u)
                    *reinterpret_cast<uint4*>(U + o) = *reinterpret_cast<const uint4*>(slab + o);
                __syncthreads();
            }
            // ---- transform (lib.rs:113-116) + correspondence search (lib.rs:118-124)
            double sx[PPT], sy[PPT], dx[PPT], dy[PPT];
            uint64_t hpart = 0;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int i = tid + k * THREADS;
                sx[k] = 0.; sy[k] = 0.; dx[k] = 0.; dy[k] = 0.;
                if (i < n) {
                    const double px = src[(size_t)i * DIM + 0], py = src[(size_t)i * DIM + 1];
                    const double pz = (DIM == 3) ? src[(size_t)i * DIM + 2] : 0.0;
                    transform_apply(T, px, py, &sx[k], &sy[k]);   // transform_xy keeps z (lib.rs:52-57)
                    int idx;
                    if (SHARED_TARGET) {
                        GridView<DIM, uint32_t, uint32_t> gv;
                        gv.x = p.target.x; gv.y = p.target.y; gv.z = p.target.z;
                        gv.idx = p.target.idx; gv.cstart = p.target.cstart; gv.g = geom;
                        NnResult<DIM, uint32_t, uint32_t> r = grid_nn_search(gv, sx[k], sy[k], pz);
                        dx[k] = gv.x[r.pos]; dy[k] = gv.y[r.pos];
                        idx = r.idx;
                    } else {
                        GridView<DIM, uint16_t, uint16_t> gv;
                        gv.x = reinterpret_cast<const double*>(U + L.nn_x);
                        gv.y = reinterpret_cast<const double*>(U + L.nn_y);
                        gv.z = reinterpret_cast<const double*>(U + L.nn_z);
                        gv.idx = reinterpret_cast<const uint16_t*>(U + L.nn_idx);
                        gv.cstart = reinterpret_cast<const uint16_t*>(U + L.nn_cell);
                        gv.g = geom;
                        NnResult<DIM, uint16_t, uint16_t> r = grid_nn_search(gv, sx[k], sy[k], pz);
                        dx[k] = gv.x[r.pos]; dy[k] = gv.y[r.pos];
                        idx = r.idx;
                    }
                    hpart += mix64(((uint64_t)(uint32_t)i << 32) | (uint64_t)(uint32_t)idx);
                    if (p.out_corr) {
                        if (p.corr_mode == 2) p.out_corr[corr_base + (uint64_t)it * (uint64_t)n + (uint64_t)i] = idx;
                        else if (p.corr_mode == 1 && it + 1 == p.max_iter) p.out_corr[corr_base + (uint64_t)i] = idx;
                    }
                }
            }
            hpart = warp_sum_u64(hpart);
            if (lane == 0) ctl->hash_part[warp] = hpart;
            __syncthreads();  // search done: the union region becomes the IRLS row arrays
            if (tid == 0) {
                uint64_t h = 0;
                for (int w = 0; w < NW; ++w) h += ctl->hash_part[w];
                H = H * 0x100000001B3ull + h;
            }

            // ---- estimate_transform (lib.rs:59-84)
            Transform Ti = transform_identity();
            double prev_error = kF64Max;
            uint32_t accepted = 0, why = kBreakCap;
            if (n < 2) {
                why = kBreakNone;  // check_input_size (lib.rs:186-189,225-228)
            } else {
                for (int inner = 0; inner < kInnerMaxIter; ++inner) {
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
                    for (int q = tid; q < 6 * 2 * kChainUnroll; q += THREADS) {
                        const int kk = q / (2 * kChainUnroll);
                        ktab[q] = kk == 0 ? Ti.r00 : kk == 1 ? Ti.r01 : kk == 2 ? Ti.r10
                                : kk == 3 ? Ti.r11 : kk == 4 ? 1.0 : 0.0;
                    }
                    __syncthreads();
                    // calc_stddevs (stats.rs:49-60): per-axis median, then median of |r - median|
                    double med[2], mad[2], sigma[2];
                    {
                        bool ev[2];
                        double elo[2], esc[2];
                        const double zero2[2] = {0., 0.};
#pragma unroll
                        for (int a = 0; a < 2; ++a) {
                            esc[a] = (double)kSelBins / (6.0 * est_sigma[a]);
                            elo[a] = est_med[a] - 3.0 * est_sigma[a];
                            ev[a] = est_sigma[a] > 0.0 && esc[a] < 1e300 && fabs(elo[a]) < 1e300;
                        }
                        hist_par = block_median2<THREADS>(sel, r2, n, 0, zero2, ev, elo, esc, hist_par, med);
#pragma unroll
                        for (int a = 0; a < 2; ++a) {
                            esc[a] = (double)kSelBins / (3.0 * est_sigma[a]);
                            elo[a] = 0.0;
                        }
                        hist_par = block_median2<THREADS>(sel, r2, n, 1, med, ev, elo, esc, hist_par, mad);
#pragma unroll
                        for (int a = 0; a < 2; ++a) {
                            sigma[a] = kPpf34 * mad[a];   // stats.rs:39-47
                            est_med[a] = med[a];
                            est_sigma[a] = sigma[a];
                        }
                    }
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
                                if (sigma[0] != 0.) cc.x = huber_drho(rr.x * rr.x) * g0;
                                if (sigma[1] != 0.) cc.y = huber_drho(rr.y * rr.y) * g1;
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
                    if (warp == 0) {
                        const double s = chain_run(chain_role(lane, r2, c2, b2, ktab), npad / kChainUnroll);
                        if (lane < 12) ctl->sums[lane] = s;
                    }
                    if (warp == (NW > 1 ? 1 : 0)) {
                        double e = 0.0;
                        for (int i = 0; i < npad; i += kChainUnroll) {
#pragma unroll
                            for (int u = 0; u < kChainUnroll; ++u) e = e + rho[i + u];
                        }
                        if (lane == 0) ctl->sums[12] = e;
                    }
                    __syncthreads();
                    // solve + break tests + update: every thread, identically (lib.rs:67-81)
                    double jtj[9], jtr[3], delta[3];
#pragma unroll
                    for (int q = 0; q < 9; ++q) jtj[q] = ctl->sums[q];
#pragma unroll
                    for (int q = 0; q < 3; ++q) jtr[q] = ctl->sums[9 + q];
                    const double error = ctl->sums[12];
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
            T = transform_mul(Ti, T);  // lib.rs:127
            if (tid == 0 && p.out_inner) p.out_inner[(size_t)pair * p.max_iter + it] = accepted | (why << 16);
            __syncthreads();  // rows are dead; next iteration restages the target
        }
        if (tid == 0) {
            p.out_T[(size_t)pair * 6 + 0] = T.r00; p.out_T[(size_t)pair * 6 + 1] = T.r01;
            p.out_T[(size_t)pair * 6 + 2] = T.r10; p.out_T[(size_t)pair * 6 + 3] = T.r11;
            p.out_T[(size_t)pair * 6 + 4] = T.tx;  p.out_T[(size_t)pair * 6 + 5] = T.ty;
            if (p.out_status) p.out_status[pair] = kPairOk;
            if (p.out_hash) p.out_hash[pair] = H;
        }
        __syncthreads();  // ctl->pair is rewritten at the top of the loop
    }
}

}  // namespace icp
