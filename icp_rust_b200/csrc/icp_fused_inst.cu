// icp_fused_inst.cu — one capacity of the fused per-pair ICP kernel (icp_fused.cuh), four
// specialisations: {2-D, 3-D} x {per-pair target in shared memory, shared target in HBM}.
#include "icp_instances.h"

#define ICP_CAT_(a, b) a##b
#define ICP_CAT(a, b) ICP_CAT_(a, b)
#define ICP_K(DIM, SHARED) \
    icp_fused_kernel<ICP_INST_THREADS, ICP_INST_PPT, DIM, SHARED, ICP_INST_MINB, (ICP_INST_ORDERED != 0)>

#if ICP_INST_ORDERED
#define ICP_L(DIM, SHARED) nullptr
#else
#define ICP_L(DIM, SHARED) icp_lean_kernel<ICP_INST_THREADS, ICP_INST_PPT, DIM, SHARED, ICP_INST_MINB, (ICP_INST_SVC != 0)>
#endif

namespace icp {
extern const Instance ICP_CAT(kFusedInstance, ICP_INST_ID) = {
    ICP_INST_THREADS, ICP_INST_PPT, ICP_INST_MINB, (ICP_INST_ORDERED != 0), (ICP_INST_SVC != 0),
    {{ICP_K(2, false), ICP_K(2, true)}, {ICP_K(3, false), ICP_K(3, true)}},
    {{ICP_L(2, false), ICP_L(2, true)}, {ICP_L(3, false), ICP_L(3, true)}}};
}  // namespace icp
