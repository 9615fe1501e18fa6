// Old (tile per thread, small_eval.cuh) against new (quad layout, small_quad.cuh) form of the fused small-N evaluator on the same
// evaluations: logL must agree bitwise, the gradient to rounding; kernel time per evaluation and SM for both.
// Usage: quad_time [N per band x3 = 60 50 40]
#include "small_quad.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
using namespace gpcc;

template <int MT, int MB>
__global__ void __launch_bounds__(MT, MB) k_old(DevProblem p, EvalBatch b, int T, int fwd) {
    extern __shared__ __align__(16) double smem[];
    const int e = blockIdx.x;
    small::eval_one<K_M32>(p, T, smem, b.delays + (size_t)e * p.L, b.alpha + (size_t)e * p.L, b.rho[e], b.want_grad != 0, fwd != 0,
                           b.ll + e, b.grad + (size_t)e * (p.L + 1), b.info + e);
}
template <int MT, int MB>
__global__ void __launch_bounds__(MT, MB) k_new(DevProblem p, EvalBatch b, int T, int fwd) {
    extern __shared__ __align__(16) double smem[];
    const int e = blockIdx.x;
    const small::QuadOwner q = small::quad_layout(T, threadIdx.x);
    small::eval_one_q<K_M32>(p, T, q, smem, b.delays + (size_t)e * p.L, b.alpha + (size_t)e * p.L, b.rho[e], b.want_grad != 0, fwd != 0,
                             b.ll + e, b.grad + (size_t)e * (p.L + 1), b.info + e);
}
int main(int argc, char** argv) {
    int nb[3] = {60, 50, 40};
    for (int i = 0; i < 3 && i + 1 < argc; ++i) nb[i] = atoi(argv[i + 1]);
    const int L = 3, N = nb[0] + nb[1] + nb[2];
    std::vector<double> t(N), r(N), s2(N, 0.5625), sb(N);
    std::vector<int> band(N);
    srand(1);
    DevProblem p;
    p.N = N; p.L = L; p.kernel_id = K_M32;
    for (int l = 0, i = 0; l < L; ++l) { p.band_start[l] = i; for (int q = 0; q < nb[l]; ++q, ++i) { band[i] = l; t[i] = 20.0 * rand() / RAND_MAX; r[i] = 2.0 * rand() / RAND_MAX - 1.0; sb[i] = 500.0 * (l + 1); } }
    p.band_start[L] = N;
    double *dt, *dr, *ds2, *dsb; int* dband;
    cudaMalloc(&dt, N * 8); cudaMalloc(&dr, N * 8); cudaMalloc(&ds2, N * 8); cudaMalloc(&dsb, N * 8); cudaMalloc(&dband, N * 4);
    cudaMemcpy(dt, t.data(), N * 8, cudaMemcpyHostToDevice); cudaMemcpy(dr, r.data(), N * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(ds2, s2.data(), N * 8, cudaMemcpyHostToDevice); cudaMemcpy(dsb, sb.data(), N * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dband, band.data(), N * 4, cudaMemcpyHostToDevice);
    p.t = dt; p.resid = dr; p.y = dr; p.s2 = ds2; p.sigb = dsb; p.band = dband;
    const int MMAX = 148 * 2 * 8;
    std::vector<double> delays(MMAX * L), alpha(MMAX * L), rho(MMAX);
    for (int m = 0; m < MMAX; ++m) { rho[m] = 1.0 + 5.0 * rand() / RAND_MAX; for (int l = 0; l < L; ++l) { delays[m * L + l] = l ? 10.0 * rand() / RAND_MAX : 0.0; alpha[m * L + l] = 0.5 + 2.0 * rand() / RAND_MAX; } }
    EvalBatch b;
    double *dd, *da, *drho, *dll, *dg; int* dinfo;
    cudaMalloc(&dd, MMAX * L * 8); cudaMalloc(&da, MMAX * L * 8); cudaMalloc(&drho, MMAX * 8); cudaMalloc(&dll, MMAX * 8); cudaMalloc(&dg, MMAX * (L + 1) * 8); cudaMalloc(&dinfo, MMAX * 4);
    cudaMemcpy(dd, delays.data(), MMAX * L * 8, cudaMemcpyHostToDevice); cudaMemcpy(da, alpha.data(), MMAX * L * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(drho, rho.data(), MMAX * 8, cudaMemcpyHostToDevice);
    b.delays = dd; b.alpha = da; b.rho = drho; b.ll = dll; b.grad = dg; b.info = dinfo;
    const int T = (N + 1 + 7) / 8, ntiles = T * (T + 1) / 2, threads = (ntiles + 31) / 32 * 32;
    if (T > small::QMAX_T) { printf("N too large for the quad layout\n"); return 1; }
    const int qthreads = (small::quad_threads(T) + 31) / 32 * 32;
    auto kold = threads <= 128 ? k_old<128, 3> : k_old<192, 2>;
    const int lthreads = threads;
    auto knew = k_new<128, 2>;
    const size_t sm_old = small::eval_smem_bytes(T, 1), sm_new = small::qeval_smem_bytes(T, 1);
    cudaFuncSetAttribute(kold, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_old);
    cudaFuncSetAttribute(knew, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_new);
    int occ_old = 0, occ_new = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_old, kold, lthreads, sm_old);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_new, knew, qthreads, sm_new);
    printf("N=%d T=%d | old: %d threads, smem %zu B, %d CTAs/SM | new: %d threads (%d used), smem %zu B, %d CTAs/SM\n", N, T, threads, sm_old,
           occ_old, qthreads, small::quad_threads(T), sm_new, occ_new);
    double *dll2, *dg2; int* dinfo2;
    cudaMalloc(&dll2, MMAX * 8); cudaMalloc(&dg2, MMAX * (L + 1) * 8); cudaMalloc(&dinfo2, MMAX * 4);
    EvalBatch b2 = b; b2.ll = dll2; b2.grad = dg2; b2.info = dinfo2;
    int rc = 0;
    for (int fwd = 0; fwd < 2; ++fwd)
        for (int M : {1, 148, 148 * 2, 148 * 2 * 8}) {
            b.M = b2.M = M; b.want_grad = b2.want_grad = fwd ? 0 : 1;
            cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
            cudaMemset(dll2, 0xff, MMAX * 8); cudaMemset(dg2, 0xff, MMAX * (L + 1) * 8);
            kold<<<M, lthreads, sm_old>>>(p, b, T, fwd);
            knew<<<M, qthreads, sm_new>>>(p, b2, T, fwd);
            cudaEventRecord(e0);
            kold<<<M, lthreads, sm_old>>>(p, b, T, fwd);
            cudaEventRecord(e1);
            knew<<<M, qthreads, sm_new>>>(p, b2, T, fwd);
            cudaEventRecord(e2);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 1; }
            float ms_old, ms_new; cudaEventElapsedTime(&ms_old, e0, e1); cudaEventElapsedTime(&ms_new, e1, e2);
            std::vector<double> l1(M), l2(M), g1(M * (L + 1)), g2(M * (L + 1));
            std::vector<int> i1(M), i2(M);
            cudaMemcpy(l1.data(), dll, M * 8, cudaMemcpyDeviceToHost); cudaMemcpy(l2.data(), dll2, M * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(g1.data(), dg, M * (L + 1) * 8, cudaMemcpyDeviceToHost); cudaMemcpy(g2.data(), dg2, M * (L + 1) * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(i1.data(), dinfo, M * 4, cudaMemcpyDeviceToHost); cudaMemcpy(i2.data(), dinfo2, M * 4, cudaMemcpyDeviceToHost);
            int nbit = 0, ninfo = 0; double gerr = 0, lerr = 0;
            for (int m = 0; m < M; ++m) {
                if (memcmp(&l1[m], &l2[m], 8) != 0) { ++nbit; lerr = std::max(lerr, fabs(l1[m] - l2[m]) / fabs(l1[m])); }
                if (i1[m] != i2[m]) ++ninfo;
                if (!fwd) for (int c = 0; c <= L; ++c) gerr = std::max(gerr, fabs(g1[m * (L + 1) + c] - g2[m * (L + 1) + c]) / (1e-300 + fabs(g1[m * (L + 1) + c])));
            }
            if (lerr > 1e-13 || ninfo || gerr > 1e-9 || !(gerr == gerr)) rc = 2;
            const int waves = (M + 147) / 148;
            printf("%s M=%5d: old %.3f ms (%.1f us per evaluation and SM), new %.3f ms (%.1f us) -> x%.2f | logL not bitwise equal: %d (max rel %.1e), info differs: %d, max rel gradient difference %.1e | ll[0] %.12g %.12g\n",
                   fwd ? "forward" : "sweep  ", M, ms_old, ms_old * 1e3 / waves, ms_new, ms_new * 1e3 / waves, ms_old / ms_new, nbit, lerr, ninfo, gerr, l1[0], l2[0]);
        }
    printf(rc ? "MISMATCH\n" : "AGREE\n");
    return rc;
}
