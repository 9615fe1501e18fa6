// Kernel-id dispatch of the fused small-N path.  small_sweep.cu and small_fit.cu are compiled once per kernel family
// (OU, rbf, matern32, matern52: src/util.jl:15-52) into separate objects so that the template instantiations build in parallel.
#include "gpcc_internal.h"
#include "kernfun.cuh"

namespace gpcc {

#define GPCC_DECL(k)                                                                                         \
    cudaError_t small_sweep_launch_k##k(const DevProblem& p, const EvalBatch& b, cudaStream_t s);           \
    cudaError_t small_fit_launch_k##k(const DevProblem& p, const FitParams& fp, const FitBuffers& fb, cudaStream_t s);
GPCC_DECL(0) GPCC_DECL(1) GPCC_DECL(2) GPCC_DECL(3)
#undef GPCC_DECL
static_assert(K_OU == 0 && K_RBF == 1 && K_M32 == 2 && K_M52 == 3, "object files are named after the kernel ids");

cudaError_t small_sweep_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t s) {
    switch (p.kernel_id) {
        case K_OU:  return small_sweep_launch_k0(p, b, s);
        case K_RBF: return small_sweep_launch_k1(p, b, s);
        case K_M32: return small_sweep_launch_k2(p, b, s);
        case K_M52: return small_sweep_launch_k3(p, b, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t small_fit_launch(const DevProblem& p, const FitParams& fp, const FitBuffers& fb, cudaStream_t s) {
    switch (p.kernel_id) {
        case K_OU:  return small_fit_launch_k0(p, fp, fb, s);
        case K_RBF: return small_fit_launch_k1(p, fp, fb, s);
        case K_M32: return small_fit_launch_k2(p, fp, fb, s);
        case K_M52: return small_fit_launch_k3(p, fp, fb, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace gpcc
