// Dependent-chain latencies on B200: DFMA, DMUL, MUFU.RCP64H, LDS.128 -> DFMA, __syncthreads.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, int iters) {
    __shared__ __align__(16) double sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = 1.0 + 1e-9 * threadIdx.x;
    __syncthreads();
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    double m = a;
    for (int i = 0; i < iters; ++i) m = m * b;
    long long t2 = clock64();
    double r = m + 2.0;
    for (int i = 0; i < iters; ++i) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r)); r = y + 2.0; }
    long long t3 = clock64();
    double s = r;
    int idx = threadIdx.x & 3;
    for (int i = 0; i < iters; ++i) { double vx, vy; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(vx), "=d"(vy) : "r"((unsigned)__cvta_generic_to_shared(&sm[2 * idx]))); s = fma(s, vx, vy); idx = (idx + (int)(s > 1e300)) & 3; }
    long long t4 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    long long t5 = clock64();
    out[threadIdx.x] = a + m + r + s;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
    for (int threads : {32, 192}) {
        const int iters = 4096;
        lat<<<1, threads>>>(out, cyc, iters); lat<<<1, threads>>>(out, cyc, iters);
        long long h[5]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %3d: DFMA chain %.1f cyc, DMUL chain %.1f, rcp.approx+DADD chain %.1f, LDS.128->DFMA chain %.1f, __syncthreads %.1f\n", threads,
               (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters);
    }
    return 0;
}
