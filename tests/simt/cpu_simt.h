// cpu_simt.h — lock-step CTA emulator for the CPU-only test suite.  TEST INFRASTRUCTURE.
//
// Lets tests/ compile the kernel headers of icp_rust_b200/csrc with g++ and execute one CTA at
// a time: every CUDA thread is a ucontext coroutine; __syncthreads and the warp collectives
// block a coroutine until its peers arrive.  Scheduling order is selectable (forward, reverse,
// seeded shuffle) so that a missing barrier shows up as a result that depends on the order.
// Nothing here is linked into or loaded by the shipped library; it is never a fallback.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

struct double2 { double x, y; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
struct dim3_ { unsigned x = 1, y = 1, z = 1; };

namespace simt {

struct Warp {
    uint32_t slots[2][32];
    int arrived = 0;
    uint32_t gen = 0;
};

struct NamedBar { int arrived = 0; uint32_t gen = 0; };

struct Cta {
    std::vector<unsigned char> smem;
    int nthreads = 0;
    int bar_arrived = 0;
    uint32_t bar_gen = 0;
    int bar_or[2] = {0, 0};
    NamedBar named[16];
    std::vector<Warp> warps;
};

struct Thread {
    ucontext_t ctx;
    std::vector<unsigned char> stack;
    bool done = false;
    dim3_ tid, bid;
};

struct State {
    Cta* cta = nullptr;
    Thread* cur = nullptr;
    ucontext_t sched;
    dim3_ block_dim, grid_dim;
    int schedule = 0;      // 0 forward, 1 reverse, >=2 shuffled with this seed
    uint64_t yields = 0;
};

inline State& st()
{
    static State s;
    return s;
}

inline void yield() { st().yields++; swapcontext(&st().cur->ctx, &st().sched); }

inline std::function<void()>*& entry_fn()
{
    static std::function<void()>* f = nullptr;
    return f;
}

inline void trampoline()
{
    (*entry_fn())();
    st().cur->done = true;
    swapcontext(&st().cur->ctx, &st().sched);
}

// Runs `body` once per thread of each CTA of a grid, CTAs one after another.
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, std::function<void()> body)
{
    State& S = st();
    S.block_dim.x = block;
    S.grid_dim.x = grid;
    entry_fn() = &body;
    for (unsigned b = 0; b < grid; ++b) {
        Cta cta;
        cta.smem.assign(smem_bytes + 64, 0xCD);  // poison: kernels must initialise what they read
        cta.nthreads = (int)block;
        cta.warps.resize((block + 31) / 32);
        S.cta = &cta;
        std::vector<Thread> threads(block);
        for (unsigned t = 0; t < block; ++t) {
            Thread& th = threads[t];
            th.stack.resize(256 * 1024);
            th.tid.x = t;
            th.bid.x = b;
            getcontext(&th.ctx);
            th.ctx.uc_stack.ss_sp = th.stack.data();
            th.ctx.uc_stack.ss_size = th.stack.size();
            th.ctx.uc_link = nullptr;
            makecontext(&th.ctx, (void (*)())trampoline, 0);
        }
        std::vector<int> order(block);
        for (unsigned t = 0; t < block; ++t) order[t] = (int)t;
        if (S.schedule == 1) std::reverse(order.begin(), order.end());
        std::mt19937 rng((uint32_t)S.schedule * 7919u + b);
        size_t alive = block;
        while (alive) {
            if (S.schedule >= 2) std::shuffle(order.begin(), order.end(), rng);
            for (int t : order) {
                Thread& th = threads[(size_t)t];
                if (th.done) continue;
                S.cur = &th;
                swapcontext(&S.sched, &th.ctx);
                if (th.done) --alive;
            }
        }
        S.cta = nullptr;
    }
}

inline int lane() { return (int)(st().cur->tid.x & 31u); }
inline Warp& warp() { return st().cta->warps[st().cur->tid.x >> 5]; }

// all 32 lanes deposit a word; returns the generation's buffer once everyone has arrived
inline const uint32_t* warp_exchange(uint32_t v)
{
    Warp& w = warp();
    const uint32_t g = w.gen;
    const int lanes = std::min(32, st().cta->nthreads - (int)(st().cur->tid.x & ~31u));
    w.slots[g & 1][lane()] = v;
    if (++w.arrived == lanes) {
        w.arrived = 0;
        w.gen = g + 1;
    } else {
        while (w.gen == g) yield();
    }
    return w.slots[g & 1];
}

}  // namespace simt

// ------------------------------------------------------------------ CUDA surface used by the kernels
#define ICP_DEV static inline
#define ICP_DEV_NOINLINE static inline
#define ICP_DEV_MEMBER inline
typedef const double2* smem_ref;
static inline smem_ref smem_ref_of(const double2* p) { return p; }
static inline smem_ref smem_adv2(smem_ref r, int n) { return r + n; }
static inline double2 smem_ld2(smem_ref r, int idx) { return r[idx]; }
#define ICP_HD static inline
#define ICP_GLOBAL static inline
#define ICP_SHARED_DYN(name) unsigned char* name = simt::st().cta->smem.data()
#define ICP_LAUNCH_BOUNDS(t, b)

#define threadIdx (simt::st().cur->tid)
#define blockIdx (simt::st().cur->bid)
#define blockDim (simt::st().block_dim)
#define gridDim (simt::st().grid_dim)

static inline void __syncthreads()
{
    simt::Cta& c = *simt::st().cta;
    const uint32_t g = c.bar_gen;
    if (++c.bar_arrived == c.nthreads) {
        c.bar_arrived = 0;
        c.bar_gen = g + 1;
    } else {
        while (c.bar_gen == g) simt::yield();
    }
}
static inline int __syncthreads_or(int pred)
{
    simt::Cta& c = *simt::st().cta;
    const uint32_t g = c.bar_gen;
    if (c.bar_arrived == 0) c.bar_or[g & 1] = 0;
    c.bar_or[g & 1] |= (pred != 0);
    if (++c.bar_arrived == c.nthreads) {
        c.bar_arrived = 0;
        c.bar_gen = g + 1;
    } else {
        while (c.bar_gen == g) simt::yield();
    }
    return c.bar_or[g & 1];
}
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_exchange(0); }

// named barriers: bar.arrive (does not wait) / bar.sync with an id and a thread count
static inline void icp_bar_arrive(int id, int count)
{
    simt::NamedBar& b = simt::st().cta->named[id];
    if (++b.arrived == count) { b.arrived = 0; b.gen++; }
}
static inline void icp_bar_sync(int id, int count)
{
    simt::NamedBar& b = simt::st().cta->named[id];
    const uint32_t g = b.gen;
    if (++b.arrived == count) { b.arrived = 0; b.gen = g + 1; }
    else while (b.gen == g) simt::yield();
}

// programmatic dependent launch: the emulator runs one launch at a time
static inline void icp_pdl_launch_dependents() {}
static inline void icp_pdl_wait() {}
static inline void __threadfence() {}

static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return simt::warp_exchange(v)[src & 31]; }
static inline int __shfl_sync(unsigned m, int v, int src) { return (int)__shfl_sync(m, (uint32_t)v, src); }
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int mask)
{
    const int l = simt::lane();
    return simt::warp_exchange(v)[(l ^ mask) & 31];
}
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d)
{
    const int l = simt::lane();
    const uint32_t* s = simt::warp_exchange(v);
    return (l >= (int)d) ? s[l - (int)d] : v;
}
static inline uint32_t __ballot_sync(unsigned, int pred)
{
    const uint32_t* s = simt::warp_exchange(pred ? 1u : 0u);
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= (s[i] & 1u) << i;
    return r;
}
static inline uint32_t __reduce_add_sync(unsigned, uint32_t v)
{
    const uint32_t* s = simt::warp_exchange(v);
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r += s[i];
    return r;
}
static inline uint32_t __reduce_max_sync(unsigned, uint32_t v)
{
    const uint32_t* s = simt::warp_exchange(v);
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r = s[i] > r ? s[i] : r;
    return r;
}
static inline int __reduce_min_sync(unsigned, int v)
{
    const uint32_t* s = simt::warp_exchange((uint32_t)v);
    int r = (int)s[0];
    for (int i = 1; i < 32; ++i) r = (int)s[i] < r ? (int)s[i] : r;
    return r;
}
static inline int __reduce_max_sync(unsigned, int v)
{
    const uint32_t* s = simt::warp_exchange((uint32_t)v);
    int r = (int)s[0];
    for (int i = 1; i < 32; ++i) r = (int)s[i] > r ? (int)s[i] : r;
    return r;
}
static inline int __any_sync(unsigned, int pred)
{
    const uint32_t* s = simt::warp_exchange(pred ? 1u : 0u);
    int r = 0;
    for (int i = 0; i < 32; ++i) r |= (int)(s[i] & 1u);
    return r;
}
// NOTE: partial warps (block size not a multiple of 32) are not supported by the collectives.

static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { auto o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { auto o = *p; if (v > o) *p = v; return o; }

static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __double2int_rd(double v) { return (int)std::floor(v); }
static inline float __double2float_rn(double v) { return (float)v; }
static inline int __float2int_rd(float v) { return (int)std::floor(v); }
using std::fmaxf;
using std::fminf;
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
using std::cbrt;
using std::fabs;
using std::floor;
using std::fmax;
using std::fmin;
using std::fmod;
using std::sqrt;
