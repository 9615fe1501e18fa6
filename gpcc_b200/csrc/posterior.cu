// Posterior over delay candidates on the device.
// Restates /root/reference/src/getprobabilities.jl:10-20: joint = loglikel .+ logprior;
// posterior = exp.(joint .- logsumexp(joint)); the 1-argument method (:1-6) uses a "prior" of ones.
// M is at most ~1e5 doubles, so a single CTA is the right size (latency bound, not a roofline kernel).
#include "gpcc_internal.h"
#include <algorithm>
#include <cmath>

namespace gpcc {
namespace {

__global__ void __launch_bounds__(1024) posterior_kernel(int M, const double* __restrict__ ll,
                                                         const double* __restrict__ logprior, double* __restrict__ out,
                                                         int joint_already) {
    __shared__ double red[32];
    __shared__ double bcast;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto joint = [&](int m) { return joint_already ? ll[m] : ll[m] + (logprior ? logprior[m] : 1.0); };
    double mx = -INFINITY;
    for (int m = tid; m < M; m += blockDim.x) mx = fmax(mx, joint(m));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        double v = -INFINITY;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v = fmax(v, red[w]);
        bcast = v;
    }
    __syncthreads();
    mx = bcast;
    double s = 0.0;
    if (isfinite(mx))
        for (int m = tid; m < M; m += blockDim.x) s += exp(joint(m) - mx);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
        bcast = mx + log(v);          // logsumexp (StatsFuns semantics: -Inf when every entry is -Inf, +Inf passes through)
    }
    __syncthreads();
    const double lse = bcast;
    for (int m = tid; m < M; m += blockDim.x) out[m] = exp(joint(m) - lse);
}

// ---- sharded form: per-device partial max / sum-exp BEFORE the all-gather ------------------------------
__global__ void __launch_bounds__(1024) posterior_partial_kernel(int per, int rec, int jcol, double* __restrict__ send) {
    __shared__ double red[32];
    __shared__ double bcast;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double mx = -INFINITY;
    for (int k = tid; k < per; k += blockDim.x) mx = fmax(mx, send[(size_t)k * rec + jcol]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        double v = -INFINITY;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v = fmax(v, red[w]);
        bcast = v;
    }
    __syncthreads();
    mx = bcast;
    double s = 0.0;
    if (isfinite(mx))
        for (int k = tid; k < per; k += blockDim.x) s += exp(send[(size_t)k * rec + jcol] - mx);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
        send[(size_t)per * rec + 0] = mx;
        send[(size_t)per * rec + 1] = v;
    }
}

// after the all-gather: logsumexp from the G pairs (max_g, sum_g), then exp(joint - logsumexp) for every candidate
__global__ void posterior_combine_kernel(int G, int per, int rec, int jcol, const double* __restrict__ recv, double* __restrict__ out) {
    const size_t blk = (size_t)per * rec + 2;
    double mx = -INFINITY;
    for (int g = 0; g < G; ++g) mx = fmax(mx, recv[g * blk + (size_t)per * rec]);
    double s = 0.0;
    if (isfinite(mx))
        for (int g = 0; g < G; ++g) {
            const double mg = recv[g * blk + (size_t)per * rec], sg = recv[g * blk + (size_t)per * rec + 1];
            if (isfinite(mg)) s += sg * exp(mg - mx);
        }
    const double lse = mx + log(s);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < G * per; q += gridDim.x * blockDim.x) {
        const int g = q / per, k = q % per;
        out[q] = exp(recv[g * blk + (size_t)k * rec + jcol] - lse);
    }
}

}  // namespace

cudaError_t posterior_partial_launch(int per, int rec, int jcol, double* d_send, cudaStream_t stream) {
    posterior_partial_kernel<<<1, 1024, 0, stream>>>(per, rec, jcol, d_send);
    return cudaGetLastError();
}

cudaError_t posterior_combine_launch(int G, int per, int rec, int jcol, const double* d_recv, double* d_out, cudaStream_t stream) {
    const int blocks = std::min(148, (G * per + 255) / 256);
    posterior_combine_kernel<<<blocks, 256, 0, stream>>>(G, per, rec, jcol, d_recv, d_out);
    return cudaGetLastError();
}

cudaError_t posterior_launch(int M, const double* d_ll, const double* d_logprior, double* d_out, cudaStream_t stream,
                             bool joint_already) {
    posterior_kernel<<<1, 1024, 0, stream>>>(M, d_ll, d_logprior, d_out, joint_already ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace gpcc
