// Yardstick only (cuSOLVER is not used by the library): cusolverDnDpotrf / cusolverDnDpotrfBatched on SPD matrices of the
// sizes of BASELINE configs 4/5, as TFLOP/s on the N^3/3 flop model -- the library kernel the tiled path is compared with.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <cusolverDn.h>
__global__ void fill_spd(double* A, int N, int batch) {
    const size_t total = (size_t)N * N * batch;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q % N), j = (int)((q / N) % N);
        const double d = (double)(i - j);
        A[q] = exp(-fabs(d) / 50.0) + (i == j ? 0.5 : 0.0);       // OU covariance + noise: SPD
    }
}
int main() {
    cusolverDnHandle_t h; cusolverDnCreate(&h);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int N : {768, 3072, 6144, 12288}) {
        const int batch = N <= 768 ? 512 : (N <= 3072 ? 64 : (N <= 6144 ? 32 : 4));
        double* A; cudaMalloc(&A, sizeof(double) * (size_t)N * N * batch);
        int lwork = 0; cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, N, A, N, &lwork);
        double* work; cudaMalloc(&work, sizeof(double) * lwork);
        int* info; cudaMalloc(&info, sizeof(int) * batch);
        for (int mode = 0; mode < 2; ++mode) {      // 0: potrf one matrix after the other, 1: potrfBatched
            if (mode == 1 && N > 3072) continue;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                fill_spd<<<1184, 256>>>(A, N, batch);
                std::vector<double*> ptr(batch); for (int b = 0; b < batch; ++b) ptr[b] = A + (size_t)b * N * N;
                double** dptr; cudaMalloc(&dptr, sizeof(double*) * batch); cudaMemcpy(dptr, ptr.data(), sizeof(double*) * batch, cudaMemcpyHostToDevice);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                if (mode == 0) for (int b = 0; b < batch; ++b) cusolverDnDpotrf(h, CUBLAS_FILL_MODE_LOWER, N, ptr[b], N, work, lwork, info + b);
                else cusolverDnDpotrfBatched(h, CUBLAS_FILL_MODE_LOWER, N, dptr, N, info, batch);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
                cudaFree(dptr);
            }
            int hinfo = -1; cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
            printf("N=%5d batch %3d %-13s: %.3f ms per matrix, %.2f TFLOP/s (N^3/3), info %d\n", N, batch, mode ? "potrfBatched" : "potrf (loop)",
                   best / batch, batch * (double)N * N * N / 3.0 / (best * 1e-3) / 1e12, hinfo);
        }
        cudaFree(A); cudaFree(work); cudaFree(info);
    }
    return 0;
}
