// placeholder (replaced below in this round): tiled large-N path
#include "gpcc_internal.h"
namespace gpcc {
cudaError_t large_eval(const DevProblem&, const EvalBatch&, LargeWorkspace&, cudaStream_t, bool, LargeTimings*) {
    return cudaErrorNotSupported;
}
void large_workspace_release(LargeWorkspace&) {}
}  // namespace gpcc
