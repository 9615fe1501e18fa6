// Microbenchmark: FP64 ceilings of one B200 -- DFMA (vector pipe) and DMMA.8x8x4 (mma.sync m8n8k4 f64).
// MEASURED_PEAKS.json has no FP64 figure; these are the roofline denominators for the Cholesky/sweep kernels.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
    double a[16];
    const double x = 1.0000001, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-6, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            const int iters = 20000, blocks = sms * (2048 / threads);
            cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 16 * iters * (double)blocks * threads;
            if (rep) printf("DFMA threads/blk %4d blocks %d : %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM @1.965GHz)\n", threads, blocks, ms, flops / ms / 1e9, flops / 2 / (ms * 1e-3) / sms / 1.965e9);
            cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            flops = 512.0 * 8 * iters * (double)blocks * threads / 32;
            if (rep) printf("DMMA threads/blk %4d blocks %d : %.3f ms  %.2f TFLOP/s\n", threads, blocks, ms, flops / ms / 1e9);
        }
    }
    printf("sms %d clock %d kHz\n", sms, prop.clockRate);
    return 0;
}
