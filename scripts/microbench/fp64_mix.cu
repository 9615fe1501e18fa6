// Do DFMA and DMMA share one FP64 datapath on B200?  Run them alone and interleaved (different warps / same warp).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: DFMA only, 1: DMMA only, 2: even warps DFMA / odd warps DMMA, 3: both in every warp
__global__ void k(double* out, int iters, int mode) {
    double f[16], c[8][2];
    const double x = 1.0000001, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1e-3;
    const bool do_f = mode == 0 || mode == 3 || (mode == 2 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_m = mode == 1 || mode == 3 || (mode == 2 && ((threadIdx.x >> 5) & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (do_f) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], x, y);
        }
        if (do_m) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, threads = 512, blocks = sms * 4, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"DFMA only", "DMMA only", "even warps DFMA, odd warps DMMA", "DFMA + DMMA in every warp"};
    for (int mode = 0; mode < 4; ++mode) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); k<<<blocks, threads>>>(out, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        double warps = (double)blocks * threads / 32;
        double wf = (mode == 0 || mode == 3) ? warps : (mode == 2 ? warps / 2 : 0);
        double wm = (mode == 1 || mode == 3) ? warps : (mode == 2 ? warps / 2 : 0);
        double flops = wf * 32 * 16 * 2.0 * iters + wm * 8 * 512.0 * iters;
        printf("%-34s %.3f ms  %.2f TFLOP/s total\n", names[mode], ms, flops / ms / 1e9);
    }
    return 0;
}
