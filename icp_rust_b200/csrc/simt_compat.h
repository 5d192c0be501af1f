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
#define ICP_HD __host__ __device__ __forceinline__
#define ICP_GLOBAL __global__
#define ICP_SHARED_DYN(name) extern __shared__ __align__(16) unsigned char name[]
#define ICP_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#endif

#define ICP_FULL_MASK 0xffffffffu
