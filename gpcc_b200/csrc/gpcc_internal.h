// Internal declarations shared by the CUDA translation units of libgpcc_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace gpcc {

constexpr int MAX_BANDS = 8;   // == GPCC_MAX_BANDS

// Problem data resident on one device (uploaded once by gpcc_problem_create).
struct DevProblem {
    int N = 0;            // total points
    int L = 0;            // bands
    int kernel_id = 0;
    const double* t = nullptr;      // [N] observation times, bands concatenated
    const double* resid = nullptr;  // [N] Y - bbar            (gpccfixdelay_marginaliseb.jl:85,98)
    const double* y = nullptr;      // [N] Y                   (:85; right-hand side of the postb solve, :250)
    const double* s2 = nullptr;     // [N] sigma^2  (Sobs)     (:89)
    const double* sigb = nullptr;   // [N] Sigma_b[band(i)]    (:94-96, B = Q Sigma_b Q')
    const int* band = nullptr;      // [N] band index of point i
    int band_start[9] = {0};        // band l occupies [band_start[l], band_start[l+1])
};

// One batch of (delay, hyper-parameter) evaluations, all arrays on the device.
struct EvalBatch {
    int M = 0;
    const double* delays = nullptr;  // [M][L]
    const double* alpha = nullptr;   // [M][L]
    const double* rho = nullptr;     // [M]
    int want_grad = 0;
    double* ll = nullptr;            // [M]
    double* grad = nullptr;          // [M][L+1]
    int* info = nullptr;             // [M]
    double* dump_kinv = nullptr;     // optional [M][N*N] dense column-major K~^-1 (both triangles)
    double* dump_a = nullptr;        // optional [M][N]   a = K~^-1 (Y - bbar)
    int mode_postb = 0;              // 1: factor Sobs + K WITHOUT B and solve against Y (postb, :248-250)
    double* dump_chol = nullptr;     // optional [M][N*N] dense column-major Cholesky factor (lower, upper part zero); tiled path,
                                     // forward mode only (the fitted state behind postb / pred, fitstate.cu)
};

// ---- small-N path: fused register-resident symmetric sweep (small_sweep.cu) -------------------
// Largest N the fused kernel handles (N+1 padded to a multiple of 8 gives <= SMALL_MAX_T tile rows).
constexpr int SMALL_TILE = 8;
constexpr int SMALL_MAX_T = 25;
inline bool small_path_supports(int N) { return (N + 1 + SMALL_TILE - 1) / SMALL_TILE <= SMALL_MAX_T; }
cudaError_t small_sweep_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t stream);
cudaError_t small_sweep_init();   // sets max dynamic shared memory attributes once per device
// blocked (eight pivots per step) form of the fused evaluator (small_block.cu)
bool small_block_supports(int N);
cudaError_t small_block_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t stream);
// fragment-layout form: tiles spread over warps, DMMA for panel and update, several matrices per CTA (small_frag.cu)
bool small_frag_supports(int N);
cudaError_t small_frag_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t stream);
// tensor-pipe variant of the fused evaluator (small_dmma.cu)
bool small_dmma_supports(int N);
cudaError_t small_dmma_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t stream);


// ---- large-N path: tiled matrix in HBM, blocked sweep with DMMA trailing updates (large_path.cu) --
struct LargeWorkspace {
    void* impl = nullptr;
};
struct LargeTimings {
    double ms_assembly = 0, ms_factor = 0, ms_gradreduce = 0;
    long long launches = 0;
};
cudaError_t large_eval(const DevProblem& p, const EvalBatch& b, LargeWorkspace& ws, cudaStream_t stream, bool profile,
                       LargeTimings* timings);
void large_workspace_release(LargeWorkspace& ws);

// ---- posterior over candidates (posterior.cu): src/getprobabilities.jl:10-20 --------------------
cudaError_t posterior_launch(int M, const double* d_ll, const double* d_logprior, double* d_out, cudaStream_t stream,
                             bool joint_already = false);

// ---- NCCL, loaded at run time so that single-GPU use needs no NCCL at all (nccl_bridge.cpp) -------
struct NcclBridge;
NcclBridge* nccl_bridge_create(const std::vector<int>& devs, std::string& err);
void nccl_bridge_destroy(NcclBridge* b);
int nccl_bridge_allgather(NcclBridge* b, const std::vector<double*>& send, const std::vector<double*>& recv, int count,
                          const std::vector<cudaStream_t>& streams, std::string& err);

}  // namespace gpcc

struct gpcc_ctx;
namespace gpcc {
int posterior_on_devices(gpcc_ctx* ctx, int M, const double* ll, const double* logprior, double* out_post);
}
