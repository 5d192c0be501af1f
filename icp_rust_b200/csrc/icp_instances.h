// icp_instances.h — table of the compiled fused-kernel instances (one translation unit each,
// icp_fused_inst.cu built with -DICP_INST_ID/THREADS/PPT/MINB/ORDERED; list in the Makefile).
#pragma once
#include "icp_lean.cuh"
#include "icp_stream.cuh"

namespace icp {

typedef void (*FusedKernel)(FusedParams);

struct Instance {
    int threads, ppt, minb;
    bool ordered;         // reference-order sums (bit-exact) vs warp-tree reductions
    bool svc;             // lean kernels: warps 0 and 1 are dedicated service warps (capacity (threads - 64) * ppt)
    FusedKernel k[2][2];  // [dim-2][shared target]
    FusedKernel lean[2][2];   // icp_lean.cuh schedule of the warp-tree mode (null for ordered instances)
    int capacity(bool lean_schedule) const { return (threads - (lean_schedule && svc ? 64 : 0)) * ppt; }
};

typedef void (*StreamKernel)(StreamParams);
constexpr int kStreamThreads = 512;
struct StreamInstance {
    int threads;
    StreamKernel k[2][2];  // [dim-2][shared target]
};
extern const StreamInstance kStreamInstance;

#define ICP_DECLARE_INSTANCE(id) extern const Instance kFusedInstance##id;
ICP_DECLARE_INSTANCE(0) ICP_DECLARE_INSTANCE(1) ICP_DECLARE_INSTANCE(2) ICP_DECLARE_INSTANCE(3)
ICP_DECLARE_INSTANCE(4) ICP_DECLARE_INSTANCE(5) ICP_DECLARE_INSTANCE(6) ICP_DECLARE_INSTANCE(7)
ICP_DECLARE_INSTANCE(8) ICP_DECLARE_INSTANCE(9) ICP_DECLARE_INSTANCE(10) ICP_DECLARE_INSTANCE(11)
ICP_DECLARE_INSTANCE(12) ICP_DECLARE_INSTANCE(13)
#undef ICP_DECLARE_INSTANCE

}  // namespace icp
