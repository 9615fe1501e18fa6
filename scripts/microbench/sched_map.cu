// Which warps of a CTA (and of two co-resident CTAs) share a scheduler (SM sub-partition)?  There is no %schedulerid, so the
// map is inferred from throughput: every warp runs the same stream of independent DFMAs; the FP64 pipe of a sub-partition
// takes one warp instruction every 2 cycles, so a warp that shares its pipe with n-1 others runs at 1/n of the lone rate.
// Phase A: all warps run together (per-warp cycles show the load of its scheduler).  Phase B: warps run in pairs (0,w): the
// pair is slow exactly when w sits on the scheduler of warp 0.  Usage: sched_map [warps per CTA = 6] [CTAs per SM = 2]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double dfma_stream(double a, int iters) {
    double r0 = a, r1 = a + 1, r2 = a + 2, r3 = a + 3, r4 = a + 4, r5 = a + 5, r6 = a + 6, r7 = a + 7;
    const double x = 1.0000001, y = 1e-9;
    for (int i = 0; i < iters; ++i) {
        r0 = fma(r0, x, y); r1 = fma(r1, x, y); r2 = fma(r2, x, y); r3 = fma(r3, x, y);
        r4 = fma(r4, x, y); r5 = fma(r5, x, y); r6 = fma(r6, x, y); r7 = fma(r7, x, y);
    }
    return r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
}

// mode 0: every warp runs the stream at the same time; mode 1: only warp 0 of CTA slot 0 and warp `other` (global index
// over the SM's resident warps: cta_slot * warps + w) run it
__global__ void k(long long* cycles, unsigned* smid, double* sink, int iters, int mode, int other, int nsm, int regs_dummy) {
    __shared__ int dummy;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned sm;
    asm("mov.u32 %0, %%smid;" : "=r"(sm));
    const int slot = blockIdx.x / nsm;            // CTAs are dealt round-robin over the SMs: blocks b and b + nsm share an SM (checked via smid)
    const int gw = slot * nw + warp;
    if (threadIdx.x == 0) dummy = 0;
    __syncthreads();
    bool run = mode == 0 || gw == 0 || gw == other;
    long long t0 = clock64();
    double r = 0;
    if (run) r = dfma_stream(1.0 + lane, iters);
    long long t1 = clock64();
    if (lane == 0) {
        cycles[blockIdx.x * nw + warp] = run ? t1 - t0 : 0;
        smid[blockIdx.x] = sm;
    }
    if (r == 123.456) sink[0] = r + dummy;
}

int main(int argc, char** argv) {
    const int nw = argc > 1 ? atoi(argv[1]) : 6, per_sm = argc > 2 ? atoi(argv[2]) : 2;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount, nb = nsm * per_sm, iters = 20000;
    long long* dc; unsigned* ds; double* dsink;
    cudaMalloc(&dc, sizeof(long long) * nb * nw); cudaMalloc(&ds, 4 * nb); cudaMalloc(&dsink, 8);
    std::vector<long long> c(nb * nw); std::vector<unsigned> s(nb);
    // enough dynamic shared memory that exactly per_sm CTAs fit on an SM
    const int smem = (int)(prop.sharedMemPerMultiprocessor / per_sm) - 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    auto launch = [&](int mode, int other) {
        k<<<nb, nw * 32, smem>>>(dc, ds, dsink, iters, mode, other, nsm, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(c.data(), dc, sizeof(long long) * nb * nw, cudaMemcpyDeviceToHost);
        cudaMemcpy(s.data(), ds, 4 * nb, cudaMemcpyDeviceToHost);
    };
    launch(0, 0); launch(0, 0);
    // pick an SM on which blocks b and b + nsm really met
    int b0 = -1;
    for (int b = 0; b < nsm && b0 < 0; ++b) { bool ok = true; for (int q = 1; q < per_sm; ++q) ok = ok && s[b + q * nsm] == s[b]; if (ok) b0 = b; }
    if (b0 < 0) { printf("no SM with the expected co-residency\n"); return 1; }
    const double lone = 8.0 * iters * 2.0;   // cycles of the stream on an otherwise idle FP64 pipe (2 per DFMA)
    printf("%d warps per CTA, %d CTAs per SM; block %d and friends on SM %u; lone stream = %.0f cycles\n", nw, per_sm, b0, s[b0], lone);
    printf("phase A (all warps run): relative time per warp = warps on its scheduler\n");
    for (int q = 0; q < per_sm; ++q) {
        printf("  CTA slot %d:", q);
        for (int w = 0; w < nw; ++w) printf(" w%d %.2f", w, c[(b0 + q * nsm) * nw + w] / lone);
        printf("\n");
    }
    printf("phase B (warp 0 of slot 0 + one other): time of warp 0 relative to lone; ~2 = same scheduler\n");
    for (int q = 0; q < per_sm; ++q) {
        printf("  CTA slot %d:", q);
        for (int w = 0; w < nw; ++w) {
            const int other = q * nw + w;
            if (other == 0) { printf(" w0 self"); continue; }
            launch(1, other);
            printf(" w%d %.2f", w, c[b0 * nw + 0] / lone);
        }
        printf("\n");
    }
    return 0;
}
