// icp_instances.h — table of the compiled fused-kernel instances (one translation unit each,
// icp_fused_inst.cu built with -DICP_INST_ID/THREADS/PPT/MINB).
#pragma once
#include "icp_fused.cuh"

namespace icp {

typedef void (*FusedKernel)(FusedParams);

struct Instance {
    int threads, ppt, minb;
    FusedKernel k[2][2];  // [dim-2][shared target]
};

extern const Instance kFusedInstance0, kFusedInstance1, kFusedInstance2, kFusedInstance3,
    kFusedInstance4, kFusedInstance5;

}  // namespace icp
