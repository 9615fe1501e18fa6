// postb and predictions on the device, reusing K~^-1 from the sweep.
//   gpcc_postb           src/gpccfixdelay_marginaliseb.jl:248-252
//   gpcc_predict         :259-307
//   gpcc_predict_loglik  :311-343
// These run once per `gpcc` call (not inside the fit loop); the kernels here are simple tiled FP64
// kernels, not the roofline-critical ones.  The reference re-factorises KSobsB with a dense `\` (LU) on
// every pred call; here K~^-1 and a = K~^-1 (Y - bbar) are computed once per (delays, alpha, rho) by the
// sweep kernel and cached on the problem.
#include "state.h"
#include "kernfun.cuh"
#include <cmath>
#include <cstring>
#include <limits>

using namespace gpcc;

namespace {

constexpr double JITTER = 1e-8;       // :69, :279
constexpr double SIGMA_FLOOR = 1e-6;  // :303
constexpr double LOG2PI = 1.8378770664093454835606594728112;

struct HyperDev {   // small by-value parameter block
    double delays[GPCC_MAX_BANDS];
    double alpha[GPCC_MAX_BANDS];
    double sigb[GPCC_MAX_BANDS];
    double mub[GPCC_MAX_BANDS];
    double rho;
    int L;
};

// kB* = delayedCovariance(kernel, alpha, tau, rho, tarray, ttest) + Q Sigma_b Q*'   (:264, :269), column-major N x NT
template <int KID>
__global__ void cross_cov_kernel(int N, int NT, const double* __restrict__ t, const int* __restrict__ band,
                                 const double* __restrict__ tt, const int* __restrict__ bandt, HyperDev h,
                                 double* __restrict__ out) {
    const KernParams kp = make_kern_params(KID, h.rho);
    const int j = blockIdx.y;
    const int bj = bandt[j];
    const double tj = tt[j] - h.delays[bj], aj = h.alpha[bj];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int bi = band[i];
        double v = (h.alpha[bi] * aj) * kern_value<KID>((t[i] - h.delays[bi]) - tj, kp);
        if (bi == bj) v += h.sigb[bi];
        out[(size_t)j * N + i] = v;
    }
}

// C[M x Nc] = op(A) * B, column-major.  TRANS_A=false: A is M x K (lda=M).  TRANS_A=true: A is K x M (lda=K), C = A' B.
template <bool TRANS_A>
__global__ void __launch_bounds__(256) gemm_kernel(int M, int Nc, int K, const double* __restrict__ A,
                                                   const double* __restrict__ B, double* __restrict__ C) {
    __shared__ double As[16][65];
    __shared__ double Bs[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            int kk, mm;
            if (TRANS_A) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
            const int gm = m0 + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < M && gk < K) v = TRANS_A ? A[(size_t)gm * K + gk] : A[(size_t)gk * M + gm];
            As[kk][mm] = v;
            const int kb = e & 15, nb = e >> 4;
            const int gn = n0 + nb, gkb = k0 + kb;
            Bs[kb][nb] = (gn < Nc && gkb < K) ? B[(size_t)gn * K + gkb] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][tx + 16 * i]; b[i] = Bs[kk][ty + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + tx + 16 * i, gn = n0 + ty + 16 * j;
            if (gm < M && gn < Nc) C[(size_t)gn * M + gm] = acc[i][j];
        }
}

// mu_j = kB*[:,j]' a + mu_b[band*(j)]   (:283-285);  q_j = kB*[:,j]' V[:,j] ; one warp per test point
__global__ void col_dots_kernel(int N, int NT, const double* __restrict__ ks, const double* __restrict__ V,
                                const double* __restrict__ a, const int* __restrict__ bandt, HyperDev h,
                                double* __restrict__ mu, double* __restrict__ q) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= NT) return;
    const int lane = threadIdx.x & 31;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < N; i += 32) {
        const double k = ks[(size_t)j * N + i];
        s1 = fma(k, a[i], s1);
        s2 = fma(k, V[(size_t)j * N + i], s2);
    }
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) { mu[j] = s1 + h.mub[bandt[j]]; q[j] = s2; }
}

// sd_j = sqrt(max(c**_jj - q_j + JITTER, 1e-6)), c**_jj = alpha^2 k(0) + Sigma_b   (:272, :279, :303)
__global__ void diag_sd_kernel(int NT, const double* __restrict__ q, const int* __restrict__ bandt, HyperDev h,
                               double* __restrict__ sd) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= NT) return;
    const int b = bandt[j];
    const double c = h.alpha[b] * h.alpha[b] + h.sigb[b];
    sd[j] = sqrt(fmax(c - q[j] + JITTER, SIGMA_FLOOR));
}

// Sigma_pred = cB - sym(kB*' V) + JITTER I   (:272-279), column-major NT x NT; optional + diag(extra) (:319)
template <int KID>
__global__ void pred_cov_kernel(int NT, const double* __restrict__ G, const double* __restrict__ tt,
                                const int* __restrict__ bandt, HyperDev h, const double* __restrict__ extra_diag,
                                double* __restrict__ S) {
    const KernParams kp = make_kern_params(KID, h.rho);
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= NT) return;
    const int bi = bandt[i], bj = bandt[j];
    double c = (h.alpha[bi] * h.alpha[bj]) * kern_value<KID>((tt[i] - h.delays[bi]) - (tt[j] - h.delays[bj]), kp);
    if (bi == bj) c += h.sigb[bi];
    double v = c - 0.5 * (G[(size_t)j * NT + i] + G[(size_t)i * NT + j]);
    if (i == j) { v += JITTER; if (extra_diag) v += extra_diag[i]; }
    S[(size_t)j * NT + i] = v;
}

// One-CTA right-looking Cholesky of an n x n column-major matrix in global memory followed by the Gaussian
// log-density of y under N(mu, S): out[0] = logpdf, info = k>0 if leading minor k is not PD.  (:325)
__global__ void __launch_bounds__(1024) chol_logpdf_kernel(int n, double* __restrict__ S, const double* __restrict__ y,
                                                           const double* __restrict__ mu, double* __restrict__ z,
                                                           double* __restrict__ out, int* __restrict__ info) {
    __shared__ double s_piv;
    __shared__ int s_info;
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_info = 0;
    for (int i = tid; i < n; i += nt) z[i] = y[i] - mu[i];
    __syncthreads();
    double logdet = 0.0;
    for (int k = 0; k < n; ++k) {
        if (tid == 0) {
            const double d = S[(size_t)k * n + k];
            if (!(d > 0.0) && s_info == 0) s_info = k + 1;
            s_piv = sqrt(d);
        }
        __syncthreads();
        if (s_info) break;
        const double piv = s_piv;
        for (int i = k + tid; i < n; i += nt) S[(size_t)k * n + i] /= piv;
        __syncthreads();
        // trailing update, lower triangle: column j (k<j<n), rows i>=j
        const int m = n - k - 1;
        for (long long e = tid; e < (long long)m * m; e += nt) {
            const int jj = (int)(e / m), ii = (int)(e % m);
            if (ii >= jj) {
                const int i = k + 1 + ii, j = k + 1 + jj;
                S[(size_t)j * n + i] -= S[(size_t)k * n + i] * S[(size_t)k * n + j];
            }
        }
        // forward substitution for z interleaved: z_k final, z_i -= L_ik z_k
        if (tid == 0) z[k] /= piv;
        __syncthreads();
        const double zk = z[k];
        for (int i = k + 1 + tid; i < n; i += nt) z[i] -= S[(size_t)k * n + i] * zk;
        logdet += 2.0 * log(piv);
        __syncthreads();
    }
    __syncthreads();
    if (s_info) { if (tid == 0) { *info = s_info; out[0] = -INFINITY; } return; }
    double q = 0.0;
    for (int i = tid; i < n; i += nt) q += z[i] * z[i];
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((tid & 31) == 0) red[tid >> 5] = q;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < (nt >> 5); ++w) s += red[w];
        out[0] = -0.5 * ((double)n * LOG2PI + logdet + s);
        *info = 0;
    }
}

// Block sums of (K+Sobs)^-1 over band pairs and band sums of (K+Sobs)^-1 Y, then the L x L algebra of :248-252.
__global__ void __launch_bounds__(256) postb_kernel(int N, int L, const double* __restrict__ Minv,
                                                    const double* __restrict__ a, const int* __restrict__ band_start,
                                                    HyperDev h, double* __restrict__ out_mu, double* __restrict__ out_S) {
    __shared__ double G[GPCC_MAX_BANDS][GPCC_MAX_BANDS];
    __shared__ double hv[GPCC_MAX_BANDS];
    __shared__ double red[8];
    const int tid = threadIdx.x;
    for (int l = 0; l < L; ++l) {
        for (int m = 0; m <= L; ++m) {      // m == L: band sum of a
            double s = 0.0;
            const int i0 = band_start[l], i1 = band_start[l + 1];
            if (m < L) {
                const int j0 = band_start[m], j1 = band_start[m + 1];
                const long long cnt = (long long)(i1 - i0) * (j1 - j0);
                for (long long e = tid; e < cnt; e += 256) {
                    const int i = i0 + (int)(e % (i1 - i0)), j = j0 + (int)(e / (i1 - i0));
                    s += Minv[(size_t)j * N + i];
                }
            } else {
                for (int i = i0 + tid; i < i1; i += 256) s += a[i];
            }
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = s;
            __syncthreads();
            if (tid == 0) {
                double v = 0.0;
                for (int w = 0; w < 8; ++w) v += red[w];
                if (m < L) G[l][m] = v; else hv[l] = v;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        // A = Sigma_b^-1 + Q'(Sobs+K)^-1 Q ; Sigma_post = A^-1 (Gauss-Jordan, A is SPD) (:248)
        double A[GPCC_MAX_BANDS][2 * GPCC_MAX_BANDS];
        for (int l = 0; l < L; ++l)
            for (int m = 0; m < L; ++m) {
                A[l][m] = 0.5 * (G[l][m] + G[m][l]) + (l == m ? 1.0 / h.sigb[l] : 0.0);
                A[l][L + m] = (l == m) ? 1.0 : 0.0;
            }
        for (int k = 0; k < L; ++k) {
            const double p = 1.0 / A[k][k];
            for (int m = 0; m < 2 * L; ++m) A[k][m] *= p;
            for (int l = 0; l < L; ++l)
                if (l != k) {
                    const double f = A[l][k];
                    for (int m = 0; m < 2 * L; ++m) A[l][m] -= f * A[k][m];
                }
        }
        // mu_post = Sigma_post ((Q'/(Sobs+K)) Y + Sigma_b \ mu_b) (:250); Sigma_post symmetrised (:252)
        for (int l = 0; l < L; ++l) {
            double s = 0.0;
            for (int m = 0; m < L; ++m) s += A[l][L + m] * (hv[m] + h.mub[m] / h.sigb[m]);
            out_mu[l] = s;
            for (int m = 0; m < L; ++m) out_S[m * L + l] = 0.5 * (A[l][L + m] + A[m][L + l]);
        }
    }
}

struct Scratch {   // RAII device allocations for one call
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <class T> cudaError_t alloc(T** out, size_t n) {
        cudaError_t e = cudaMalloc(out, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
};

HyperDev make_hyper(const gpcc_problem* p, const double* delays, const double* alpha, double rho) {
    HyperDev h{};
    h.L = p->L; h.rho = rho;
    for (int l = 0; l < p->L; ++l) { h.delays[l] = delays[l]; h.alpha[l] = alpha[l]; h.sigb[l] = p->Sigmab[l]; h.mub[l] = p->mub[l]; }
    return h;
}

int check_hyper(const gpcc_problem* p, const double* delays, const double* alpha, double rho) {
    if (!p || !delays || !alpha) return fail(-1, "NULL argument");
    for (int l = 0; l < p->L; ++l) {
        if (!(alpha[l] > 0) || !std::isfinite(alpha[l])) return fail(-2, "all(scale .> 0) violated (delayedCovariance.jl:3)");
        if (!std::isfinite(delays[l])) return fail(-3, "non-finite delay");
    }
    if (!(rho > 0) || !std::isfinite(rho)) return fail(-4, "rho is <= 0 (delayedCovariance.jl:5-7)");
    return 0;
}

// K~^-1 (dense, column-major, on device 0) and a = K~^-1 (Y - bbar) for the given hyper-parameters.
// mode_postb: K+Sobs without B and a = (K+Sobs)^-1 Y instead (:248-250).
int inverse_on_device(gpcc_problem* p, const double* delays, const double* alpha, double rho, int mode_postb,
                      double* d_kinv, double* d_a, int* info) {
    const int L = p->L;
    int rc = reserve_eval(p, 0, 1);
    if (rc) return rc;
    EvalSlot& s = p->ctx->ds[0].slot[0];
    std::memcpy(s.delays.h, delays, L * sizeof(double));
    std::memcpy(s.alpha.h, alpha, L * sizeof(double));
    s.rho.h[0] = rho;
    rc = evaluate_on_device(p, 0, 1, 1, d_kinv, d_a, mode_postb);
    if (rc) return rc;
    *info = s.info.h[0];
    return 0;
}

template <int KID>
int predict_impl(gpcc_problem* p, const HyperDev& h, int NT, const int* d_bandt, const double* d_tt, const double* d_kinv,
                 const double* d_a, double* d_mu, double* d_sd, double* d_S, const double* d_extra, Scratch& sc,
                 cudaStream_t st) {
    const int N = p->N;
    const auto& dp = p->pd[0].dp;
    double *d_ks = nullptr, *d_V = nullptr, *d_q = nullptr, *d_G = nullptr;
    CUDA_TRY(sc.alloc(&d_ks, (size_t)N * NT));
    CUDA_TRY(sc.alloc(&d_V, (size_t)N * NT));
    CUDA_TRY(sc.alloc(&d_q, NT));
    cross_cov_kernel<KID><<<dim3((N + 255) / 256, NT), 256, 0, st>>>(N, NT, dp.t, dp.band, d_tt, d_bandt, h, d_ks);
    gemm_kernel<false><<<dim3((N + 63) / 64, (NT + 63) / 64), 256, 0, st>>>(N, NT, N, d_kinv, d_ks, d_V);   // KSobsB \ kB*
    col_dots_kernel<<<(NT + 7) / 8, 256, 0, st>>>(N, NT, d_ks, d_V, d_a, d_bandt, h, d_mu, d_q);
    if (d_sd) diag_sd_kernel<<<(NT + 255) / 256, 256, 0, st>>>(NT, d_q, d_bandt, h, d_sd);
    if (d_S) {
        CUDA_TRY(sc.alloc(&d_G, (size_t)NT * NT));
        gemm_kernel<true><<<dim3((NT + 63) / 64, (NT + 63) / 64), 256, 0, st>>>(NT, NT, N, d_ks, d_V, d_G);   // kB*' (KSobsB \ kB*)
        pred_cov_kernel<KID><<<dim3((NT + 255) / 256, NT), 256, 0, st>>>(NT, d_G, d_tt, d_bandt, h, d_extra, d_S);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int predict_common(gpcc_problem* p, const double* delays, const double* alpha, double rho, const int* ntest_per_band,
                   const double* ttest, const double* ytest, const double* sigmatest, double* out_mu, double* out_sd,
                   double* out_Sigma, double* out_ll, int* out_info) {
    int rc = check_hyper(p, delays, alpha, rho);
    if (rc) return rc;
    if (!ntest_per_band) return fail(-1, "NULL argument");
    const int L = p->L, N = p->N;
    int NT = 0;
    for (int l = 0; l < L; ++l) { if (ntest_per_band[l] < 0) return fail(-5, "negative test count"); NT += ntest_per_band[l]; }
    if (NT == 0) { if (out_ll) *out_ll = 0.0; if (out_info) *out_info = 0; return 0; }
    if (!ttest) return fail(-1, "NULL argument");
    std::vector<int> bandt(NT);
    for (int l = 0, k = 0; l < L; ++l) for (int i = 0; i < ntest_per_band[l]; ++i) bandt[k++] = l;
    DeviceState& s = p->ctx->ds[0];
    CUDA_TRY(cudaSetDevice(s.dev));
    Scratch sc;
    double *d_kinv, *d_a, *d_tt, *d_mu, *d_sd = nullptr, *d_S = nullptr, *d_extra = nullptr;
    int* d_bandt;
    CUDA_TRY(sc.alloc(&d_kinv, (size_t)N * N));
    CUDA_TRY(sc.alloc(&d_a, N));
    int info = 0;
    rc = inverse_on_device(p, delays, alpha, rho, 0, d_kinv, d_a, &info);
    if (rc) return rc;
    if (info != 0) return fail(-6, "K + Sobs + B is not positive definite for these hyper-parameters");
    CUDA_TRY(sc.alloc(&d_tt, NT));
    CUDA_TRY(sc.alloc(&d_bandt, NT));
    CUDA_TRY(sc.alloc(&d_mu, NT));
    CUDA_TRY(cudaMemcpyAsync(d_tt, ttest, NT * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    CUDA_TRY(cudaMemcpyAsync(d_bandt, bandt.data(), NT * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    const bool want_ll = out_ll != nullptr;
    if (out_sd) CUDA_TRY(sc.alloc(&d_sd, NT));
    if (out_Sigma || want_ll) CUDA_TRY(sc.alloc(&d_S, (size_t)NT * NT));
    if (want_ll) {
        std::vector<double> s2(NT);
        for (int i = 0; i < NT; ++i) s2[i] = sigmatest[i] * sigmatest[i];          // Sobs* (:317)
        CUDA_TRY(sc.alloc(&d_extra, NT));
        CUDA_TRY(cudaMemcpyAsync(d_extra, s2.data(), NT * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
    }
    const HyperDev h = make_hyper(p, delays, alpha, rho);
    switch (p->kernel_id) {
        case K_OU:  rc = predict_impl<K_OU>(p, h, NT, d_bandt, d_tt, d_kinv, d_a, d_mu, d_sd, d_S, d_extra, sc, s.stream); break;
        case K_RBF: rc = predict_impl<K_RBF>(p, h, NT, d_bandt, d_tt, d_kinv, d_a, d_mu, d_sd, d_S, d_extra, sc, s.stream); break;
        case K_M32: rc = predict_impl<K_M32>(p, h, NT, d_bandt, d_tt, d_kinv, d_a, d_mu, d_sd, d_S, d_extra, sc, s.stream); break;
        default:    rc = predict_impl<K_M52>(p, h, NT, d_bandt, d_tt, d_kinv, d_a, d_mu, d_sd, d_S, d_extra, sc, s.stream); break;
    }
    if (rc) return rc;
    if (out_mu) CUDA_TRY(cudaMemcpyAsync(out_mu, d_mu, NT * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (out_sd) CUDA_TRY(cudaMemcpyAsync(out_sd, d_sd, NT * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (out_Sigma) CUDA_TRY(cudaMemcpyAsync(out_Sigma, d_S, (size_t)NT * NT * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (want_ll) {
        double *d_y, *d_z, *d_out;
        int* d_info;
        CUDA_TRY(sc.alloc(&d_y, NT));
        CUDA_TRY(sc.alloc(&d_z, NT));
        CUDA_TRY(sc.alloc(&d_out, 1));
        CUDA_TRY(sc.alloc(&d_info, 1));
        CUDA_TRY(cudaMemcpyAsync(d_y, ytest, NT * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        chol_logpdf_kernel<<<1, 1024, 0, s.stream>>>(NT, d_S, d_y, d_mu, d_z, d_out, d_info);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out_ll, d_out, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        int hinfo = 0;
        CUDA_TRY(cudaMemcpyAsync(&hinfo, d_info, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        if (out_info) *out_info = hinfo;
    }
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return 0;
}

}  // namespace

extern "C" {

int gpcc_postb(gpcc_problem* p, const double* delays, const double* alpha, double rho, double* out_mu, double* out_Sigma) {
    int rc = check_hyper(p, delays, alpha, rho);
    if (rc) return rc;
    if (!out_mu || !out_Sigma) return fail(-1, "NULL argument");
    const int L = p->L, N = p->N;
    DeviceState& s = p->ctx->ds[0];
    CUDA_TRY(cudaSetDevice(s.dev));
    Scratch sc;
    double *d_minv, *d_a, *d_mu, *d_S;
    int* d_bs;
    CUDA_TRY(sc.alloc(&d_minv, (size_t)N * N));
    CUDA_TRY(sc.alloc(&d_a, N));
    CUDA_TRY(sc.alloc(&d_mu, L));
    CUDA_TRY(sc.alloc(&d_S, L * L));
    CUDA_TRY(sc.alloc(&d_bs, L + 1));
    int info = 0;
    rc = inverse_on_device(p, delays, alpha, rho, 1, d_minv, d_a, &info);     // Sobs + K, without B (:248)
    if (rc) return rc;
    if (info != 0) return fail(-6, "K + Sobs is not positive definite for these hyper-parameters");
    CUDA_TRY(cudaMemcpyAsync(d_bs, p->band_start.data(), (L + 1) * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    postb_kernel<<<1, 256, 0, s.stream>>>(N, L, d_minv, d_a, d_bs, make_hyper(p, delays, alpha, rho), d_mu, d_S);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_mu, d_mu, L * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaMemcpyAsync(out_Sigma, d_S, L * L * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return 0;
}

int gpcc_predict(gpcc_problem* p, const double* delays, const double* alpha, double rho, const int* ntest_per_band,
                 const double* ttest, double* out_mu, double* out_sd, double* out_Sigma) {
    return predict_common(p, delays, alpha, rho, ntest_per_band, ttest, nullptr, nullptr, out_mu, out_sd, out_Sigma,
                          nullptr, nullptr);
}

int gpcc_predict_loglik(gpcc_problem* p, const double* delays, const double* alpha, double rho, const int* ntest_per_band,
                        const double* ttest, const double* ytest, const double* sigmatest, double* out_ll, int* out_info) {
    if (!ytest || !sigmatest || !out_ll) return fail(-1, "NULL argument");
    return predict_common(p, delays, alpha, rho, ntest_per_band, ttest, ytest, sigmatest, nullptr, nullptr, nullptr, out_ll,
                          out_info);
}

}  // extern "C"
