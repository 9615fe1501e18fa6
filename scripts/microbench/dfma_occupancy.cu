// DFMA throughput of one SM as a function of resident warps and per-thread ILP (1 CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters) {
    double a[ILP];
    const double x = 1.0000001;
    double y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = 1e-9 * (i + 1 + threadIdx.x);
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(-y[i & 7], x + y[(i >> 3) & 7], a[i]);   // 3 distinct operands like the sweep
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(double* out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps : {1, 2, 4, 6, 7, 8, 12, 16, 24, 32}) {
        const int iters = 4000;
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); k<ILP><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        double fma_per_clk = (double)ILP * iters * warps * 32 / (ms * 1e-3) / 1.965e9;
        printf("ILP %2d warps/SM %2d : %.3f ms  %.1f FMA/clk/SM\n", ILP, warps, ms, fma_per_clk);
    }
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double* out; cudaMalloc(&out, sizeof(double) * prop.multiProcessorCount * 1024);
    run<64>(out, prop.multiProcessorCount);
    run<16>(out, prop.multiProcessorCount);
    return 0;
}
