// icp_stream_inst.cu — instances of the streaming (large-scan) ICP kernel (icp_stream.cuh).
#include "icp_instances.h"

namespace icp {
extern const StreamInstance kStreamInstance = {
    kStreamThreads,
    {{icp_stream_kernel<kStreamThreads, 2, false>, icp_stream_kernel<kStreamThreads, 2, true>},
     {icp_stream_kernel<kStreamThreads, 3, false>, icp_stream_kernel<kStreamThreads, 3, true>}}};
}  // namespace icp
