// Instruction-cache behaviour on B200: cycles per instruction of a straight-line loop body as a function of its size,
// for 1 warp and for several warps per SM executing the same code (FP64 FMA bodies: the sweep kernels' instruction mix).
#include <cstdio>
#include <cuda_runtime.h>
template <int NI>
__global__ void k(double* out, long long* cyc, int iters) {
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NI / 8; ++i) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NI>
void run(double* out, long long* cyc) {
    for (int warps : {1, 4, 12}) {
        const int iters = 64;
        k<NI><<<1, 32 * warps>>>(out, cyc, iters); k<NI><<<1, 32 * warps>>>(out, cyc, iters);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("body %6d instr (%4d KB) warps %2d: %.2f cycles/instr/warp\n", NI, NI * 16 / 1024, warps, (double)h / iters / NI);
    }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 64);
    run<64>(out, cyc); run<256>(out, cyc); run<512>(out, cyc); run<1024>(out, cyc); run<2048>(out, cyc); run<4096>(out, cyc); run<8192>(out, cyc); run<16384>(out, cyc);
    return 0;
}
