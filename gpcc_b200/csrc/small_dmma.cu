// Fused small-N evaluator, tensor-pipe variant: one CTA per (delay candidate, hyper-parameter) pair, the bordered
// matrix in registers in DMMA accumulator-fragment layout, eliminated eight pivots at a time.
//
// Same reference code as small_sweep.cu (objective of /root/reference/src/gpccfixdelay_marginaliseb.jl:133-141 with
// delayedCovariance.jl:1-38 fused in, plus the analytic gradient).  Differences in how the sweep is organised:
//   * the lower triangle is cut into 8x8 micro tiles; every WARP owns up to 32 tiles and every lane holds elements
//     (row lane/4, cols 2*(lane%4), +1) of each of them = the mma.sync.m8n8k4.f64 (DMMA.8x8x4) accumulator fragment.
//     Ownership is per warp, so every test "is this tile in the pivot row / column" is warp uniform: no divergence;
//   * block step k (pivots 8k..8k+7):  gather column k of the symmetric matrix into shared memory (panel P, pivot tile
//     D)  | barrier |  warp 0: D = L L' (8 lanes, shuffles)  | barrier |  one thread per matrix row: X = P L^-T and
//     Q = X L^-1 = P D^-1 by two 8-step substitutions (identity rows for the pivot tile give D^-1 itself)  | barrier |
//     every warp: A_ij -= X_i X_j' with TWO DMMAs per tile from 2+2 LDS.64, then the tiles of column/row k take their
//     final values Q and the pivot tile -D^-1.
//     Three barriers per EIGHT pivots instead of one per pivot, no per-pivot scaling multiplies, and the forward part is
//     a genuine blocked Cholesky (X = P L^-T), so log-det and the quadratic form keep Cholesky-grade accuracy;
//   * the matrix is padded with identity pivots to a multiple of 8; the right-hand side r = Y - bbar is carried as a
//     vector in shared memory through the same block steps (z = L^-1 r_k, quad += z'z, r_i -= X_i z, r_k <- L^-T z), so
//     after the last step it holds a = K~^-1 r.
#include "gpcc_internal.h"
#include "kernfun.cuh"
#include <climits>
#include <cmath>
#include <cstdlib>

namespace gpcc {
namespace {

constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr int SLOTS = 32;     // tiles per warp

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double bsum(double v, double* red, int tid, int nthreads) {
    v = wsum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = (nthreads + 31) >> 5;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}
// A/B-fragment layout of an 8x8 block X (rows r, k index c): [c/4][r][c%4]  -> lane t reads X[t/4][4h + t%4] at h*32 + t
__device__ __forceinline__ int afrag_off(int r, int c) { return ((c >> 2) << 5) + (r << 2) + (c & 3); }

// Branch-free reciprocal square root of a positive, well-scaled double (hardware approximation + Newton steps).
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double hd = 0.5 * d;
#pragma unroll
    for (int it = 0; it < 3; ++it) y = fma(y, fma(-hd * y, y, 0.5), y);     // y <- y (1.5 - 0.5 d y^2)
    return y;
}

// Cold paths of the slot loops, kept out of line so that the hot loop stays small (instruction cache).
__device__ __forceinline__ void gather_tile(int ti, int tj, int k, int lane, double v0, double v1, double* Pb, double* Db) {
    const int lr = lane >> 2, lc = (lane & 3) * 2;
    if (tj == k && ti > k) {
        *reinterpret_cast<double2*>(Pb + ti * 64 + afrag_off(lr, lc)) = make_double2(v0, v1);
    } else if (ti == k && tj < k) {                      // transposed: P_tj[r = col][c = row]
        Pb[tj * 64 + afrag_off(lc, lr)] = v0;
        Pb[tj * 64 + afrag_off(lc + 1, lr)] = v1;
    } else {                                             // pivot tile: full symmetric copy from the lower part
        if (lc <= lr) { Db[lr * 8 + lc] = v0; Db[lc * 8 + lr] = v0; }
        if (lc + 1 <= lr) { Db[lr * 8 + lc + 1] = v1; Db[(lc + 1) * 8 + lr] = v1; }
    }
}
__device__ __forceinline__ double2 writeback_tile(int ti, int tj, int k, int lane, const double* Qb) {
    const int lr = lane >> 2, lc = (lane & 3) * 2;
    if (tj == k && ti > k) return *reinterpret_cast<const double2*>(Qb + ti * 64 + 2 * lane);
    if (ti == k && tj < k) return make_double2(Qb[tj * 64 + lc * 8 + lr], Qb[tj * 64 + (lc + 1) * 8 + lr]);
    const double2 v = *reinterpret_cast<const double2*>(Qb + k * 64 + 2 * lane);
    return make_double2(-v.x, -v.y);
}

#define COPY_CHUNK_TO_ACC(C)                                                                               \
    _Pragma("unroll") for (int u = 0; u < 8; ++u) {                                                       \
        const double2 v_ = st[u * 32 + lane];                                                             \
        acc[(C) * 8 + u][0] = v_.x; acc[(C) * 8 + u][1] = v_.y;                                           \
    }
#define COPY_CHUNK_FROM_ACC(C)                                                                             \
    _Pragma("unroll") for (int u = 0; u < 8; ++u) st[u * 32 + lane] = make_double2(acc[(C) * 8 + u][0], acc[(C) * 8 + u][1]);

template <int KID, int MAXTHREADS, int MINBLOCKS>
__global__ void __launch_bounds__(MAXTHREADS, MINBLOCKS)
dmma_sweep_kernel(DevProblem p, EvalBatch b, int T /* tile rows = ceil(N/8) */) {
    extern __shared__ __align__(16) double smem[];
    const int N = p.N, L = p.L;
    const int Np = 8 * T;
    const int e = blockIdx.x;
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = T * (T + 1) / 2;
    const int lr = lane >> 2, lc = (lane & 3) * 2;   // this lane's row and first column inside a tile
    const int nwarps = nthreads >> 5;

    double* tsh = smem;                 // [Np] shifted times
    double* av = tsh + Np;              // [Np] alpha per point (0 on padding)
    double* sbv = av + Np;              // [Np] Sigma_b per point
    double* dadd = sbv + Np;            // [Np] sigma^2
    double* rv = dadd + Np;             // [Np] residual, later a = K~^-1 r
    double* piv = rv + Np;              // [Np] Schur pivots
    double* Pb = piv + Np;              // [T][64] panel tiles, A-fragment layout
    double* Xb = Pb + T * 64;           // [T][64] X = P L^-T, A-fragment layout
    double* Qb = Xb + T * 64;           // [T][64] Q = P D^-1, row-major (C-fragment) layout
    double* Xn = Qb + T * 64;           // [T][64] -X, A-fragment layout (the A operand of the update)
    double* Db = Xn + T * 64;           // [64] pivot tile (full symmetric), then L (lower) row-major
    double* Li = Db + 64;               // [8] 1 / L_jj
    double* rkn = Li + 8;               // [8] new r_k, [8] z = L^-1 r_k
    double* red = rkn + 16;             // [64]
    double2* stage = reinterpret_cast<double2*>(red + 64);          // [nwarps][8][32] fragment staging (assembly / gradient)
    double* part = reinterpret_cast<double*>(stage + nwarps * 256); // [T][T][8] gradient partial row sums (gradient only)
    double* partd = part + (b.want_grad ? T * T * 8 : 0);           // [T][8] column sums of the diagonal tiles
    int* bandv = reinterpret_cast<int*>(partd + (b.want_grad ? T * 8 : 0));   // [Np]
    double2* st = stage + warp * 256;

    const double rho = b.rho[e];
    const KernParams kp = make_kern_params(KID, rho);

    for (int i = tid; i < Np; i += nthreads) {
        double ts = 0.0, al = 0.0, sb = 0.0, dd = 0.0, r = 0.0;
        int bi = -1 - i;
        if (i < N) {
            bi = p.band[i];
            ts = p.t[i] - b.delays[(size_t)e * L + bi];        // delayedCovariance.jl:27
            al = b.alpha[(size_t)e * L + bi];
            sb = b.mode_postb ? 0.0 : p.sigb[i];
            dd = p.s2[i];
            r = b.mode_postb ? p.y[i] : p.resid[i];
        }
        tsh[i] = ts; av[i] = al; sbv[i] = sb; dadd[i] = dd; rv[i] = r; bandv[i] = bi; piv[i] = 1.0;
    }
    // first tile of this warp (tiles are numbered row-major over the lower triangle)
    const int q0 = warp * SLOTS;
    int ti0 = (int)((sqrtf(8.0f * (float)q0 + 1.0f) - 1.0f) * 0.5f);
    while (ti0 * (ti0 + 1) / 2 > q0) --ti0;
    while ((ti0 + 1) * (ti0 + 2) / 2 <= q0) ++ti0;
    const int tj0 = q0 - ti0 * (ti0 + 1) / 2;
    __syncthreads();

    // ---- assembly, eight tiles at a time through the staging buffer (keeps the exp code out of the unrolled part) -----
    double acc[SLOTS][2];
    {
        int ti = ti0, tj = tj0;
        for (int c = 0; c < SLOTS / 8; ++c) {
            for (int u = 0; u < 8; ++u) {
                double v0 = 0.0, v1 = 0.0;
                if (q0 + c * 8 + u < ntiles) {
                    const int i = ti * 8 + lr, j = tj * 8 + lc;
                    const double ai = av[i], ti_ = tsh[i];
                    v0 = (ai * av[j]) * kern_value<KID>(ti_ - tsh[j], kp);          // 0 when either index is padding
                    v1 = (ai * av[j + 1]) * kern_value<KID>(ti_ - tsh[j + 1], kp);
                    if (i == j) v0 += dadd[i];                                       // + Sobs
                    if (i == j + 1) v1 += dadd[i];
                    if (bandv[i] == bandv[j]) v0 += sbv[i];                          // + B
                    if (bandv[i] == bandv[j + 1]) v1 += sbv[i];
                    if (i >= N && i == j) v0 = 1.0;                                  // identity pivots on the padding
                    if (i >= N && i == j + 1) v1 = 1.0;
                }
                st[u * 32 + lane] = make_double2(v0, v1);
                if (++tj > ti) { tj = 0; ++ti; }
            }
            __syncwarp();
            switch (c) {
                case 0: COPY_CHUNK_TO_ACC(0) break;
                case 1: COPY_CHUNK_TO_ACC(1) break;
                case 2: COPY_CHUNK_TO_ACC(2) break;
                default: COPY_CHUNK_TO_ACC(3) break;
            }
            __syncwarp();
        }
    }
    double quad = 0.0;   // accumulated by thread 0
    const int nvalid = min(SLOTS, ntiles - q0);          // this warp's tile count (<= 0 for surplus warps: none exist)
    {   // gather column 0 for the first block step
        int ti = ti0, tj = tj0;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (s < nvalid && tj == 0) gather_tile(ti, tj, 0, lane, acc[s][0], acc[s][1], Pb, Db);
            if (++tj > ti) { tj = 0; ++ti; }
        }
    }

    // ---- block sweep --------------------------------------------------------------------------------------------------
    for (int k = 0; k < T; ++k) {
        // (a) column k of the symmetric matrix was gathered into Pb / Db by the scan pass of the previous step
        __syncthreads();
        // (b) warp 0: Cholesky of the 8x8 pivot tile (lane i < 8 owns row i), then z = L^-1 r_k
        if (warp == 0) {
            double row[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) row[c] = Db[(lane & 7) * 8 + c];
            double rk = rv[k * 8 + (lane & 7)];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double d = __shfl_sync(0xffffffffu, row[j], j);        // Schur pivot
                const double inv = fast_rsqrt(d);
                const double lij = (lane == j) ? d * inv : row[j] * inv;     // L[lane][j] for lane >= j
                row[j] = lij;
                if (lane == j) { piv[k * 8 + j] = d; Li[j] = inv; }
#pragma unroll
                for (int c = j + 1; c < 8; ++c) {
                    const double lcj = __shfl_sync(0xffffffffu, lij, c);     // L[c][j]
                    row[c] = fma(-lij, lcj, row[c]);                        // only lanes >= c keep a meaningful value
                }
                // forward substitution for z alongside: z_j = r_j / L_jj, r_i -= L_ij z_j
                const double zj = __shfl_sync(0xffffffffu, rk, j) * inv;
                if (lane == j) rk = zj;
                else if (lane > j) rk = fma(-lij, zj, rk);                   // lanes < j already hold finished z values
            }
            if (lane < 8) {
#pragma unroll
                for (int c = 0; c < 8; ++c) Db[lane * 8 + c] = (c <= lane) ? row[c] : 0.0;
                rkn[8 + lane] = rk;                                          // z
            }
        }
        __syncthreads();
        // (c) one thread per matrix row: X = P L^-T (forward substitution), Q = X L^-1 (backward substitution); the
        //     right-hand side goes through the same step: r_i -= X_i z, r_k <- D^-1 r_k, quad += z'z
        {
            if (tid == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) quad = fma(rkn[8 + c], rkn[8 + c], quad);
            }
            for (int i = tid; i < Np; i += nthreads) {
                const int tr = i >> 3, r = i & 7;
                double x[8];
                if (tr == k) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) x[c] = (c == r) ? 1.0 : 0.0;      // identity rows: Q becomes D^-1
                } else {
                    const double4 p0 = *reinterpret_cast<const double4*>(Pb + tr * 64 + r * 4);
                    const double4 p1 = *reinterpret_cast<const double4*>(Pb + tr * 64 + 32 + r * 4);
                    x[0] = p0.x; x[1] = p0.y; x[2] = p0.z; x[3] = p0.w; x[4] = p1.x; x[5] = p1.y; x[6] = p1.z; x[7] = p1.w;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
#pragma unroll
                    for (int q = 0; q < c; ++q) x[c] = fma(-x[q], Db[c * 8 + q], x[c]);
                    x[c] *= Li[c];
                }
                if (tr != k) {
                    *reinterpret_cast<double4*>(Xb + tr * 64 + r * 4) = make_double4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<double4*>(Xb + tr * 64 + 32 + r * 4) = make_double4(x[4], x[5], x[6], x[7]);
                    *reinterpret_cast<double4*>(Xn + tr * 64 + r * 4) = make_double4(-x[0], -x[1], -x[2], -x[3]);
                    *reinterpret_cast<double4*>(Xn + tr * 64 + 32 + r * 4) = make_double4(-x[4], -x[5], -x[6], -x[7]);
                    double dot = 0.0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) dot = fma(x[c], rkn[8 + c], dot);
                    rv[i] -= dot;
                }
#pragma unroll
                for (int c = 7; c >= 0; --c) {
#pragma unroll
                    for (int q = c + 1; q < 8; ++q) x[c] = fma(-x[q], Db[q * 8 + c], x[c]);
                    x[c] *= Li[c];
                }
                *reinterpret_cast<double4*>(Qb + tr * 64 + r * 8) = make_double4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<double4*>(Qb + tr * 64 + r * 8 + 4) = make_double4(x[4], x[5], x[6], x[7]);
                if (tr == k) {           // x is now row r of D^-1: (D^-1 r_k)[r] = x . r_k  (r_k still the old values in rv)
                    double dot = 0.0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) dot = fma(x[c], rv[k * 8 + c], dot);
                    rkn[r] = dot;
                }
            }
        }
        __syncthreads();
        if (tid < 8) rv[k * 8 + tid] = rkn[tid];     // nobody reads r_k until the next block step's barrier
        // (d) trailing update on the tensor pipe: two DMMAs per tile, every tile of the warp, no tests in the loop (tiles
        //     of row / column k are overwritten by the scan pass below; X of tile row k is stale but finite)
        {
#pragma unroll
            for (int half = 0; half < 2; ++half) {      // two passes: the two DMMAs of a tile are dependent, 32 tiles are not
                const double* ap = Xn + ti0 * 64 + half * 32 + lane;
                const double* bp = Xb + tj0 * 64 + half * 32 + lane;
                int rem = ti0 - tj0, trow = ti0;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    if (s < nvalid) dmma884(acc[s][0], acc[s][1], ap[0], bp[0]);
                    if (rem == 0) { ap += 64; bp = Xb + half * 32 + lane; rem = ++trow; } else { bp += 64; --rem; }
                }
            }
        }
        // (e) scan pass: final values of the tiles in row / column k, and gather of column k+1 for the next step
        {
            int ti = ti0, tj = tj0;
            const int k1 = k + 1;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                if (s < nvalid && (ti == k || tj == k || ti == k1 || tj == k1)) {
                    if (ti == k || tj == k) {
                        const double2 v = writeback_tile(ti, tj, k, lane, Qb);
                        acc[s][0] = v.x; acc[s][1] = v.y;
                    }
                    if ((ti == k1 || tj == k1) && k1 < T) gather_tile(ti, tj, k1, lane, acc[s][0], acc[s][1], Pb, Db);
                }
                if (++tj > ti) { tj = 0; ++ti; }
            }
        }
        // (the gather wrote Pb / Db, last read in phase (c) of this step: ordered by the barrier before (d))
    }
    __syncthreads();

    // ---- log-determinant, info, quadratic form -------------------------------------------------------------------------
    double ld = 0.0;
    int bad = INT_MAX;
    for (int i = tid; i < N; i += nthreads) {
        const double d = piv[i];
        if (!(d > 0.0)) bad = min(bad, i + 1); else ld += log(d);
    }
    ld = bsum(ld, red, tid, nthreads);
    __shared__ int s_bad;
    if (tid == 0) { s_bad = INT_MAX; red[32] = quad; }
    __syncthreads();
    if (bad != INT_MAX) atomicMin(&s_bad, bad);
    __syncthreads();
    const int info = (s_bad == INT_MAX) ? 0 : s_bad;
    quad = red[32];
    const double ll = -0.5 * ((double)N * LOG2PI + ld + quad);
    if (tid == 0) {
        b.ll[e] = info ? -INFINITY : ll;
        if (b.info) b.info[e] = info;
    }
    if (!b.want_grad) return;
    if (info) {
        if (tid <= L) b.grad[(size_t)e * (L + 1) + tid] = 0.0;
        return;
    }
    if (b.dump_a) for (int i = tid; i < N; i += nthreads) b.dump_a[(size_t)e * N + i] = rv[i];

    // ---- gradient: W = a a' - K~^-1 contracted with K and dK/drho, eight tiles at a time through the staging buffer ----
    double es = 0.0;
    {
        int ti = ti0, tj = tj0;
        for (int c = 0; c < SLOTS / 8; ++c) {
            switch (c) {
                case 0: COPY_CHUNK_FROM_ACC(0) break;
                case 1: COPY_CHUNK_FROM_ACC(1) break;
                case 2: COPY_CHUNK_FROM_ACC(2) break;
                default: COPY_CHUNK_FROM_ACC(3) break;
            }
            __syncwarp();
            for (int u = 0; u < 8; ++u) {
                if (q0 + c * 8 + u < ntiles) {
                    const double2 ainv = st[u * 32 + lane];              // -(K~^-1) entries of this lane
                    const double am[2] = {ainv.x, ainv.y};
                    const int i = ti * 8 + lr;
                    double ct[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = tj * 8 + lc + h;
                        const double W = fma(rv[i], rv[j], am[h]);        // a_i a_j - (K~^-1)_ij
                        double kv, dkv;
                        kern_value_drho<KID>(tsh[i] - tsh[j], kp, kv, dkv);
                        const double aa = av[i] * av[j];
                        double cc = W * (aa * kv), d = W * (aa * dkv);
                        if (ti == tj && j >= i) { if (j > i) cc = 0.0; d = 0.0; }     // upper part unused, dk(0) = 0
                        ct[h] = cc;
                        es += d;
                        if (b.dump_kinv && i < N && j < N && j <= i) {
                            double* out = b.dump_kinv + (size_t)e * N * N;
                            out[(size_t)j * N + i] = -am[h];
                            out[(size_t)i * N + j] = -am[h];
                        }
                    }
                    // diagonal elements count once (as a row contribution); strictly-lower elements as row and column
                    const bool dg0 = (ti == tj) && (tj * 8 + lc == i), dg1 = (ti == tj) && (tj * 8 + lc + 1 == i);
                    double rs = ct[0] + ct[1];
                    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
                    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
                    double c0 = dg0 ? 0.0 : ct[0], c1 = dg1 ? 0.0 : ct[1];
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); }
                    if ((lane & 3) == 0) part[(ti * T + tj) * 8 + lr] = rs;
                    if (lane < 4) {
                        double* dst = (ti == tj) ? (partd + ti * 8) : (part + (tj * T + ti) * 8);
                        dst[lc] = c0; dst[lc + 1] = c1;
                    }
                }
                if (++tj > ti) { tj = 0; ++ti; }
            }
            __syncwarp();
        }
    }
    es = bsum(es, red, tid, nthreads);     // also orders `part`
    double* srow = Pb;                      // reuse
    for (int i = tid; i < N; i += nthreads) {
        const int ti = i >> 3, r = i & 7;
        double s = partd[ti * 8 + r];
        for (int src = 0; src < T; ++src) s += part[(ti * T + src) * 8 + r];
        srow[i] = s;
    }
    __syncthreads();
    for (int pb = warp; pb < L; pb += nwarps) {
        double s = 0.0;
        for (int i = p.band_start[pb] + lane; i < p.band_start[pb + 1]; i += 32) s += srow[i];
        s = wsum(s);
        if (lane == 0) b.grad[(size_t)e * (L + 1) + pb] = s / b.alpha[(size_t)e * L + pb];
    }
    if (tid == 0) b.grad[(size_t)e * (L + 1) + L] = es;
}

size_t dmma_smem_bytes(int T, int want_grad) {
    const int Np = 8 * T;
    const int ntiles = T * (T + 1) / 2;
    const int warps = (ntiles + SLOTS - 1) / SLOTS;
    size_t doubles = (size_t)Np * 6 + (size_t)T * 64 * 4 + 64 + 8 + 16 + 64 + (size_t)warps * 512 +
                     (want_grad ? (size_t)T * T * 8 + (size_t)T * 8 : 0);
    return doubles * sizeof(double) + (size_t)Np * sizeof(int) + 32;
}

template <int KID>
cudaError_t launch_dmma(const DevProblem& p, const EvalBatch& b, int T, cudaStream_t s) {
    const int ntiles = T * (T + 1) / 2;
    const int warps = (ntiles + SLOTS - 1) / SLOTS;
    const int threads = warps * 32;
    const size_t sm = dmma_smem_bytes(T, b.want_grad);
    static const int one_cta = getenv("GPCC_SMALL_DMMA_1CTA") ? 1 : 0;
    if (threads <= 192 && !one_cta) {
        auto kfn = dmma_sweep_kernel<KID, 192, 2>;
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dmma_smem_bytes(19, 1));
        kfn<<<b.M, threads, sm, s>>>(p, b, T);
    } else if (threads <= 256) {
        auto kfn = dmma_sweep_kernel<KID, 256, 1>;
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dmma_smem_bytes(22, 1));
        kfn<<<b.M, threads, sm, s>>>(p, b, T);
    } else {
        auto kfn = dmma_sweep_kernel<KID, 384, 1>;
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dmma_smem_bytes(27, 1));
        kfn<<<b.M, threads, sm, s>>>(p, b, T);
    }
    return cudaGetLastError();
}

}  // namespace

bool small_dmma_supports(int N) {
    const int T = (N + 7) / 8;
    return T <= 27;        // 378 tiles = 12 warps
}

cudaError_t small_dmma_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t s) {
    const int T = (p.N + 7) / 8;
    switch (p.kernel_id) {
        case K_OU:  return launch_dmma<K_OU>(p, b, T, s);
        case K_RBF: return launch_dmma<K_RBF>(p, b, T, s);
        case K_M32: return launch_dmma<K_M32>(p, b, T, s);
        case K_M52: return launch_dmma<K_M52>(p, b, T, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace gpcc
