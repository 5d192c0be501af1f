// simt_compat.h — the one place where the kernel sources see their execution environment.
//
// Product build (nvcc, sm_100a): everything below maps to the native CUDA built-ins.
// tests/simt/ additionally compiles the same kernel headers with g++ against a lock-step
// CTA emulator (ICP_SIMT_EMULATE) so that barrier placement, shared-memory races and the
// selection / reduction logic can be checked by the CPU-only test suite and ThreadSanitizer.
// The emulator is test infrastructure: nothing in the shipped library links or loads it.
#pragma once

#ifdef ICP_SIMT_EMULATE
#include "cpu_simt.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define ICP_DEV __device__ __forceinline__
#define ICP_DEV_NOINLINE static __device__ __noinline__
#define ICP_DEV_MEMBER __device__ __forceinline__
#define ICP_HD __host__ __device__ __forceinline__
#define ICP_GLOBAL __global__
#define ICP_SHARED_DYN(name) extern __shared__ __align__(128) unsigned char name[]
#define ICP_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
// Shared-memory reference to a double2 array that keeps its address space across an out-of-line
// call (a plain pointer would degrade to generic loads there).
typedef uint32_t smem_ref;
static __device__ __forceinline__ smem_ref smem_ref_of(const double2* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ smem_ref smem_adv2(smem_ref r, int n) { return r + 16u * (uint32_t)n; }
static __device__ __forceinline__ double2 smem_ld2(smem_ref r, int idx)
{
    double2 v;
    // volatile + "memory": the load must not be hoisted across the shared stores / barriers around it
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(r + 16u * (uint32_t)idx) : "memory");
    return v;
}
// Named barriers (PTX bar.sync / bar.arrive with an id and a thread count): `arrive` does not wait.
// Writes a thread made before arriving are visible to the threads that leave the matching sync.
static __device__ __forceinline__ void icp_bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
static __device__ __forceinline__ void icp_bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// Programmatic dependent launch (griddepcontrol): `launch_dependents` lets the next kernel of the
// stream start filling SM resources this grid's CTAs release; `wait` blocks the calling thread until
// the grids this launch depends on have completed and their writes are visible.  Both are no-ops in
// a launch without the programmatic-serialization attribute.
static __device__ __forceinline__ void icp_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static __device__ __forceinline__ void icp_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

#define ICP_FULL_MASK 0xffffffffu
