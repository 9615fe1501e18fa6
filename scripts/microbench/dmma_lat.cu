// Dependent / independent DMMA.8x8x4 latency on B200, and the latency of the pieces of the blocked sweep's P3 phase.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc, int iters) {
    __shared__ __align__(16) double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-3 * (i % 37);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double a = 1.0 + 1e-9 * lane, b = 1e-3;
    double c0 = 0, c1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) dmma(c0, c1, a, b);                 // accumulator-dependent chain
    long long t1 = clock64();
    double d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    for (int i = 0; i < iters; ++i) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); dmma(e0, e1, a, b); dmma(f0, f1, a, b); }   // 4 independent
    long long t2 = clock64();
    double z0 = 1e-3, z1 = 2e-3;
    for (int i = 0; i < iters; ++i) { double x0 = 0, x1 = 0; dmma(x0, x1, z0, b); z0 = x0; z1 = x1; }   // A-operand-dependent chain
    long long t3 = clock64();
    double s = 0;
    for (int i = 0; i < iters; ++i) { s += __shfl_xor_sync(0xffffffffu, s + 1.0, 1); }
    long long t4 = clock64();
    for (int i = 0; i < iters; ++i) { sm[(lane * 33 + i) & 4095] = s; s += sm[(lane * 7 + i) & 4095]; }   // STS + dependent LDS
    long long t5 = clock64();
    out[threadIdx.x] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1 + z0 + z1 + s;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
    for (int threads : {32, 192}) {
        const int iters = 2048;
        k<<<1, threads>>>(out, cyc, iters); k<<<1, threads>>>(out, cyc, iters);
        long long h[5]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %3d: DMMA acc-chain %.1f cyc, 4 independent DMMA %.1f cyc per group, DMMA A-operand chain %.1f, shfl+dadd chain %.1f, STS+LDS+DADD chain %.1f\n", threads,
               (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters);
    }
    return 0;
}
