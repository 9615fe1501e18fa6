// Per-step timeline of the fused small-N evaluator (gpcc_b200/csrc/small_eval.cuh built with -DGPCC_STEP_PROF): every warp
// stamps clock() at (0) step entry, (1) after its tile update, (2) after `publish`, (3) after the barrier; the stamps live in
// shared memory and one chosen CTA dumps them.  Usage: step_prof [N per band x3 = 60 50 40] ; prints the mean cycles per phase
// for a CTA that is alone on the GPU and for one of two co-resident CTAs at full occupancy, full sweep and forward-only.
#include "small_eval.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
using namespace gpcc;

template <int MT, int MB>
__global__ void __launch_bounds__(MT, MB) k_eval(DevProblem p, EvalBatch b, int T, int fwd) {
    extern __shared__ __align__(16) double smem[];
    const int e = blockIdx.x;
    small::eval_one<K_M32>(p, T, smem, b.delays + (size_t)e * p.L, b.alpha + (size_t)e * p.L, b.rho[e], b.want_grad != 0, fwd != 0,
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
    const int T = (N + 1 + 7) / 8, ntiles = T * (T + 1) / 2, threads = (ntiles + 31) / 32 * 32, nw = threads / 32;
#ifdef GPCC_STEP_PROF
    unsigned* dprof; const int PN = 4 * 8 * GPCC_PROF_MAXSTEPS;
    cudaMalloc(&dprof, PN * 4);
    cudaMemcpyToSymbol(small::gpcc_prof_out, &dprof, sizeof(dprof));
#else
    const int PN = 0;
#endif
    auto kfn = threads <= 128 ? k_eval<128, 3> : k_eval<192, 2>;
    const size_t sm = small::eval_smem_bytes(T, 1);
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    printf("N=%d T=%d threads=%d (%d warps), smem %zu B + %d B stamps\n", N, T, threads, nw, sm, PN * 4);
    for (int fwd = 0; fwd < 2; ++fwd)
        for (int M : {1, 148, 148 * 2, 148 * 3, 148 * 2 * 8}) {
            b.M = M; b.want_grad = fwd ? 0 : 1;
            int pb = M / 2;
#ifdef GPCC_STEP_PROF
            cudaMemcpyToSymbol(small::gpcc_prof_block, &pb, sizeof(int));
#endif
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            kfn<<<M, threads, sm>>>(p, b, T, fwd);
            cudaEventRecord(e0);
            kfn<<<M, threads, sm>>>(p, b, T, fwd);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
#ifndef GPCC_STEP_PROF
            printf("%s M=%5d: kernel %.3f ms (%.1f us per evaluation and SM)\n", fwd ? "forward" : "sweep  ", M, ms, ms * 1e3 / ((M + 147) / 148));
            (void)pb; (void)nw;
            continue;
#else
            std::vector<unsigned> h(PN);
            cudaMemcpy(h.data(), dprof, PN * 4, cudaMemcpyDeviceToHost);
            auto at = [&](int slot, int w, int k) { return h[(slot * 8 + w) * GPCC_PROF_MAXSTEPS + k]; };
            // mean over steps 8..N-9 and over warps
            double upd = 0, pub = 0, bar = 0, period = 0, skew = 0; int cnt = 0;
            for (int k = 8; k < N - 9; ++k) {
                unsigned first = ~0u, last = 0;
                for (int w = 0; w < nw; ++w) {
                    upd += (double)(at(1, w, k) - at(0, w, k)); pub += (double)(at(2, w, k) - at(1, w, k)); bar += (double)(at(3, w, k) - at(2, w, k));
                    period += (double)(at(0, w, k + 1) - at(0, w, k));
                    first = std::min(first, at(2, w, k) - at(0, 0, k)); last = std::max(last, at(2, w, k) - at(0, 0, k));
                    ++cnt;
                }
                skew += (double)(last - first);
            }
            printf("%s M=%5d: kernel %.3f ms (%.1f us per evaluation and SM) | block %d, cycles per step: period %.0f = update %.0f + publish %.0f + barrier wait %.0f ; arrival skew between warps %.0f\n",
                   fwd ? "forward" : "sweep  ", M, ms, ms * 1e3 / ((M + 147) / 148), pb, period / cnt, upd / cnt, pub / cnt, bar / cnt, skew / (N - 17));
            if (M == 148 * 2 && !fwd) {
                printf("   per-warp detail at step 40..43 (entry, +update, +publish, +barrier):\n");
                for (int k = 40; k < 44; ++k) for (int w = 0; w < nw; ++w)
                    printf("   k=%d w=%d: %u %u %u %u\n", k, w, at(0, w, k) - at(0, 0, 40), at(1, w, k) - at(0, w, k), at(2, w, k) - at(1, w, k), at(3, w, k) - at(2, w, k));
            }
#endif
        }
    return 0;
}
