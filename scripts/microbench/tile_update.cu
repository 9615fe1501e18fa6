// Microbenchmark of the sweep's inner update: 8x8 FP64 register tile, rank-1 updates from shared memory.
// Variants: barrier per step or not; loads vs no loads.  One CTA per SM, W warps.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>   // 0: loads + barrier per step, 1: loads, no barrier, 2: no loads (register operands), no barrier
__global__ void __launch_bounds__(256) k(double* out, int steps) {
    __shared__ __align__(16) double buf[2][512];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) (&buf[0][0])[i] = 1e-3 * (i % 37);
    __syncthreads();
    double A[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) A[r][c] = r + c + threadIdx.x;
    const int ti = threadIdx.x & 15, tj = (threadIdx.x >> 4) & 15;
    double xr[8], vr[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { xr[c] = 1e-3 * (c + ti); vr[c] = 1e-3 * (c + tj); }
    for (int s = 0; s < steps; ++s) {
        const double* cb = buf[s & 1];
        double v[8];
        if (MODE < 2) {
#pragma unroll
            for (int p = 0; p < 4; ++p) { const double2 t = *reinterpret_cast<const double2*>(cb + p * 64 + 2 * tj); v[2 * p] = t.x; v[2 * p + 1] = t.y; }
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = vr[c];
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            double2 x;
            if (MODE < 2) x = *reinterpret_cast<const double2*>(cb + 256 + p * 64 + 2 * ti);
            else { x.x = xr[2 * p]; x.y = xr[2 * p + 1]; }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                A[2 * p][c] = fma(-x.x, v[c], A[2 * p][c]);
                A[2 * p + 1][c] = fma(-x.y, v[c], A[2 * p + 1][c]);
            }
        }
        if (MODE == 0) __syncthreads();
    }
    double sum = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) sum += A[r][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
template <int MODE>
void run(double* out, int sms, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps : {1, 2, 4, 6, 8}) {
        const int steps = 20000;
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); k<MODE><<<sms, warps * 32>>>(out, steps); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("%-28s warps/SM %d: %.1f cycles/step (64 DFMA/thread/step) -> %.1f FMA/clk/SM\n", name, warps, ms * 1e-3 * 1.965e9 / steps,
               64.0 * warps * 32 * steps / (ms * 1e-3 * 1.965e9));
    }
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double* out; cudaMalloc(&out, sizeof(double) * prop.multiProcessorCount * 256);
    run<0>(out, prop.multiProcessorCount, "LDS + barrier per step");
    run<1>(out, prop.multiProcessorCount, "LDS, no barrier");
    run<2>(out, prop.multiProcessorCount, "register operands");
    return 0;
}
