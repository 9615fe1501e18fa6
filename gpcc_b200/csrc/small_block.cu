// Fused small-N evaluator, blocked form: eight pivots per block step, FP64 DFMA pipe, matrix in registers.
//
// Same reference code as small_sweep.cu: the objective
//   K = delayedCovariance(kernel, alpha, tau, rho, tarray) + Sobs + B ; logpdf(MvNormal(bbar, K), Y)
// (/root/reference/src/gpccfixdelay_marginaliseb.jl:133-141, src/delayedCovariance.jl:1-38) plus the analytic
// gradient 0.5 tr((a a' - K^-1) dK/dtheta).  What changes is how the symmetric sweep is scheduled:
//
//   * the lower triangle of K~ (padded with identity pivots to a multiple of 8) lives in registers, one 8x8 tile per
//     thread, exactly as in small_sweep.cu; the right-hand side r = Y - bbar is a vector in shared memory;
//   * block step k (pivots 8k..8k+7) of the symmetric sweep  [D C'; C R] -> [-D^-1, D^-1 C'; C D^-1, R - C D^-1 C']:
//       P1  the owners of tile column / row k publish the panel C to shared memory; the owner of the pivot tile
//           factors D = L L' in its own registers and publishes W = L^-1 (explicit inverse of the TRIANGULAR factor:
//           backward stable, unlike D^-1), the Schur pivots (log-det, LAPACK-style info) and zr = W r_k;
//       P3  one thread per matrix row i:  Z_i = C_i W'  (so that Z Z' = C D^-1 C'),  X_i = Z_i W = C_i D^-1,
//           r_i -= Z_i . zr;  identity rows stand in for the pivot tile, so its X rows are D^-1 itself;
//       P4  every tile:  A_ij -= Z_i Z_j'  as eight rank-1 updates from shared memory with NO barrier in between
//           (64 LDS.128 + 512 DFMA per thread); then the owners read their final values X (and -D^-1) back.
//     Two barriers per EIGHT pivots (the rank-1 kernel has one per pivot), no division or reciprocal on the per-pivot
//     critical path of 190 threads, and ONE loop body for every k (the rank-1 kernel is specialised eight-fold on the
//     pivot's position inside its tile, which overflows the instruction cache: 19 % of its stall samples are no_inst);
//   * the serial part of a block step (P1 + P3, ~1.4 k cycles) is hidden by the other matrix resident on the SM.
//   Work: T * (512 DFMA * T(T+1)/2 tiles) = N^3/2 DFMA = N^3 flop per logL+grad evaluation, as before.
#include "gpcc_internal.h"
#include "kernfun.cuh"
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstdio>

namespace gpcc {

#ifdef GPCC_BLOCK_PROF
// Dev build only: per-warp timestamps of every block step, kept in shared memory (one STS each) and dumped at exit.
__device__ long long g_block_tl[32 * 8 * 12];
__device__ int g_block_dbg;
#define DBG(bit) (g_block_dbg & (bit))
#define PROF_TL(slot) do { if ((tid & 31) == 0) tlbuf[(k * 8 + (slot)) * 12 + (tid >> 5) + 6 * gid] = clock64(); } while (0)
#else
#define PROF_TL(slot)
#define DBG(bit) 0
#endif

namespace {

constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr int CS = 64;          // doubles per row-pair part of the chunk layout (>= 2 * max tile rows)
constexpr int VLEN = 4 * CS;    // doubles per chunk-layout vector
constexpr int MAX_T = 25;       // 8 * 25 = 200 rows (N <= 199 as in small_sweep.cu)

// chunk layout: [pair-of-rows part (4)][tile][2] so that the eight values of a tile are four 16-byte loads at
// "per-thread base + immediate"
__device__ __forceinline__ int cidx(int i) { return ((i & 7) >> 1) * CS + ((i >> 3) << 1) + (i & 1); }
__device__ __forceinline__ void load8(const double* buf, int tile, double (&out)[8]) {
#pragma unroll
    for (int part = 0; part < 4; ++part) {
        const double2 v = *reinterpret_cast<const double2*>(buf + part * CS + 2 * tile);
        out[2 * part] = v.x;
        out[2 * part + 1] = v.y;
    }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void group_sync(int gid, int gthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(gthreads) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Deterministic sum over the threads of one group: xor-tree inside each warp, then the warp totals in order.
__device__ __forceinline__ double group_sum(double v, double* red, int tid, int gthreads, int gid) {
    v = warp_sum(v);
    group_sync(gid, gthreads);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    group_sync(gid, gthreads);
    double s = 0.0;
    const int nw = gthreads >> 5;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

// Reciprocal square root of a positive double: hardware seed (2^-26) and two Newton steps (~1.5 ulp); a non-positive
// pivot gives NaN, which is what flags the matrix as not positive definite.
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double e = fma(-h * y, y, 0.5);
        y = fma(y, e, y);
    }
    return y;
}

// Pivot-tile work of block step k, done by the ONE thread that owns tile (k,k): in-place Cholesky of the 8x8 block
// (strict lower part <- L, diagonal <- 1/L_jj), W = L^-1 into the (otherwise unused) upper triangle, then
// W, the pivots and zr = W r_k go to shared memory.
__device__ __forceinline__ void pivot_tile(double (&A)[8][8], int k, double* piv, double* Wb, double* zr, double* quad,
                                           const double* rv) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double d = A[j][j];
        piv[k * 8 + j] = d;
        const double inv = fast_rsqrt(d);
        A[j][j] = inv;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) A[i][j] *= inv;
#pragma unroll
        for (int c = j + 1; c < 8; ++c)
#pragma unroll
            for (int i = c; i < 8; ++i) A[i][c] = fma(-A[i][j], A[c][j], A[i][c]);
    }
    // W[i][j] (i > j) = -(1/L_ii) sum_{q=j}^{i-1} L[i][q] W[q][j], W[j][j] = 1/L_jj;  W[i][j] is kept in A[j][i]
#pragma unroll
    for (int i = 1; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) {
            double s = A[i][j] * A[j][j];
#pragma unroll
            for (int q = j + 1; q < i; ++q) s = fma(A[i][q], A[j][q], s);
            A[j][i] = -s * A[i][i];
        }
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = rv[k * 8 + q];
    double qs = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double s = A[c][c] * r[c];
#pragma unroll
        for (int q = 0; q < c; ++q) s = fma(A[q][c], r[q], s);
        zr[c] = s;
        qs = fma(s, s, qs);
#pragma unroll
        for (int q = 0; q <= c; ++q) {
            const double w = (q == c) ? A[c][c] : A[q][c];
            Wb[c * 8 + q] = w;          // W  row major (zeros above the diagonal, written once at start)
            Wb[64 + q * 8 + c] = w;     // W' row major
        }
    }
    *quad += qs;
}

template <int KID, int MAXTHREADS, int MINBLOCKS, int NMAT>
__global__ void __launch_bounds__(MAXTHREADS, MINBLOCKS)
small_block_kernel(DevProblem p, EvalBatch b, int T, int gthreads, int smem_doubles_per_group) {
    extern __shared__ __align__(16) double smem_all[];
    const int N = p.N, L = p.L;
    const int Np = T * 8;
    const int PS = T * 9;   // slots (16 B) per column-pair plane of the panel buffer: tile t, row r at 9 t + r
    const int gid = (NMAT == 1) ? 0 : (int)threadIdx.x / gthreads;
    const int tid = (NMAT == 1) ? (int)threadIdx.x : (int)threadIdx.x - gid * gthreads;
    const int e = blockIdx.x * NMAT + gid;
    if (e >= b.M) return;   // whole group leaves: its named barrier is never used
    double* smem = smem_all + (size_t)gid * smem_doubles_per_group;
    const int ntiles = T * (T + 1) / 2;
    const bool active = tid < ntiles;
    const int q = active ? tid : ntiles - 1;
    int ti = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
    while (ti * (ti + 1) / 2 > q) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
    const int tj = q - ti * (ti + 1) / 2;

    double* tsh = smem;             // shifted times t_i - tau_band(i)      (chunk layout)
    double* av = tsh + VLEN;        // alpha_band(i), 0 on padding          (chunk layout)
    double* abuf = av + VLEN;       // a = K~^-1 r for the gradient         (chunk layout)
    double* sbv = abuf + VLEN;      // Sigma_b[band(i)]                     (natural)
    double* dadd = sbv + VLEN;      // sigma_i^2                            (natural)
    double* rv = dadd + VLEN;       // r, swept along with the matrix       (natural)
    double* piv = rv + VLEN;        // Schur pivots                         (natural)
    double* Wb = piv + VLEN;        // [8][8] W = L^-1, row major, zeros above the diagonal
    double* zr = Wb + 128;          // [8] W r_k   (Wb + 64: W' row major)
    double* misc = zr + 8;          // [0] quadratic form accumulator, [1] reduction scratch, [2] info (as int)
    double* red = misc + 8;         // [16] warp totals
    double* Zb = red + 16;          // [8][VLEN] Z, one chunk-layout vector per pivot of the block
    double2* Cb = reinterpret_cast<double2*>(Zb + 8 * VLEN);   // [2][4][9 T] panel C, then X (double2 = column pair; 9 slots per tile: conflict-free owners)
    double* part = Zb;              // [T][T][8] gradient partial row sums, after the sweep (aliases Zb / Cb)
    const size_t zc_doubles = (size_t)8 * VLEN + (size_t)2 * 4 * (9 * T) * 2;
    const size_t part_doubles = b.want_grad ? (size_t)T * T * 8 : 0;
    int* bandv = reinterpret_cast<int*>(Zb + (zc_doubles > part_doubles ? zc_doubles : part_doubles));   // [Np]

#ifdef GPCC_BLOCK_PROF
    long long* tlbuf = reinterpret_cast<long long*>(smem_all + (size_t)NMAT * smem_doubles_per_group);
#endif
    const double rho = b.rho[e];
    const KernParams kp = make_kern_params(KID, rho);

    for (int i = tid; i < Np; i += gthreads) {
        const int ci = cidx(i);
        if (i < N) {
            const int bi = p.band[i];
            tsh[ci] = p.t[i] - b.delays[(size_t)e * L + bi];   // delayedCovariance.jl:27 (x - delays[l])
            av[ci] = b.alpha[(size_t)e * L + bi];
            sbv[i] = b.mode_postb ? 0.0 : p.sigb[i];
            dadd[i] = p.s2[i];
            rv[i] = b.mode_postb ? p.y[i] : p.resid[i];
            bandv[i] = bi;
        } else {
            tsh[ci] = 0.0; av[ci] = 0.0; sbv[i] = 0.0; dadd[i] = 1.0; rv[i] = 0.0;   // identity pivots on the padding
            bandv[i] = -1 - i;
        }
    }
    for (int i = tid; i < 128; i += gthreads) Wb[i] = 0.0;
    if (tid == 0) misc[0] = 0.0;
    group_sync(gid, gthreads);

    // ---- assembly of the tile in registers ------------------------------------------------------------------------
    double A[8][8];
    {
        double tc[8], ac[8];
        int bc[8];
        load8(tsh, tj, tc);
        load8(av, tj, ac);
#pragma unroll
        for (int c = 0; c < 8; ++c) bc[c] = bandv[tj * 8 + c];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = ti * 8 + r;
            const int ci = cidx(i);
            const double tr = tsh[ci], ar = av[ci], sbr = sbv[i], dr = dadd[i];
            const int br = bandv[i];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int j = tj * 8 + c;
                const double kv = kern_value<KID>(tr - tc[c], kp);
                double val = (ar * ac[c]) * kv;          // scale[l]*scale[m]*kernel  (delayedCovariance.jl:27)
                if (i == j) val += dr;                   // + Sobs                   (gpccfixdelay_marginaliseb.jl:135)
                if (br == bc[c]) val += sbr;             // + B = Q Sigma_b Q'
                A[r][c] = val;
            }
        }
    }

    // ---- block sweep ----------------------------------------------------------------------------------------------
    // Two matrices per CTA: the second one starts its sweep when the first enters its first trailing update, so that
    // the serial part of one matrix's block step (P1 + P3) runs under the other's DFMA stream from then on.
    if (NMAT == 2 && gid == 1) asm volatile("bar.sync 3, %0;" ::"r"(2 * gthreads) : "memory");
    for (int k = 0; k < T; ++k) {
        PROF_TL(0);
        const bool colo = active && (tj == k) && (ti > k);
        const bool rowo = active && (ti == k) && (tj < k);
        const bool dgo = active && (ti == k) && (tj == k);
        double2* Cw = Cb + (size_t)(k & 1) * 4 * PS;
        // P1: publish the panel; factor the pivot tile
        if (DBG(64) && !dgo) {
        } else if (colo) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int slot = ti * 9 + r;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) Cw[pp * PS + slot] = make_double2(A[r][2 * pp], A[r][2 * pp + 1]);
            }
        } else if (rowo) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int slot = tj * 9 + c;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) Cw[pp * PS + slot] = make_double2(A[2 * pp][c], A[2 * pp + 1][c]);
            }
        } else if (dgo) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {   // identity rows: their X rows become D^-1
                const int slot = k * 9 + r;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp)
                    Cw[pp * PS + slot] = make_double2((r == 2 * pp) ? 1.0 : 0.0, (r == 2 * pp + 1) ? 1.0 : 0.0);
            }
            if (!DBG(8)) pivot_tile(A, k, piv, Wb, zr, misc, rv);
        }
        PROF_TL(1);
        group_sync(gid, gthreads);
        PROF_TL(2);

        // P3: Z = C W', X = Z W, r -= Z zr on the FP64 tensor pipe (DMMA.8x8x4), one 8-row tile per warp at a time.
        // The contraction index is permuted so that every fragment is what the lane already holds: with m = lane % 4
        // the two k-halves of a DMMA pair take columns 2m and 2m+1, i.e. one 16-byte load per operand and the Z
        // accumulator feeds the second product without any shuffle.  (Doing this with one thread per row costs 4 k
        // cycles: with the 128 tile registers live there is no room to keep more than one broadcast load in flight.)
        {
            const int lane = tid & 31, row = lane >> 2, m = lane & 3;
            const int w = tid >> 5, nw = gthreads >> 5;
            const double2 wz = *reinterpret_cast<const double2*>(Wb + row * 8 + 2 * m);        // W[n=row][2m, 2m+1]
            const double2 wx = *reinterpret_cast<const double2*>(Wb + 64 + row * 8 + 2 * m);   // W[2m, 2m+1][n=row]
            const double2 zq = *reinterpret_cast<const double2*>(zr + 2 * m);
            for (int t0 = w; t0 < T; t0 += 4 * nw) {   // four independent tiles in flight per warp
                double2* cp = Cw + m * PS + row + 9 * t0;
                double* zp = Zb + (2 * m) * VLEN + (row >> 1) * CS + (row & 1) + 2 * t0;
                double* rp = rv + row + 8 * t0;
                double2 c[4];
                double z0[4], z1[4], x0[4], x1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) c[u] = (t0 + u * nw < T) ? cp[9 * u * nw] : make_double2(0.0, 0.0);
#pragma unroll
                for (int u = 0; u < 4; ++u) { z0[u] = 0.0; z1[u] = 0.0; dmma884(z0[u], z1[u], c[u].x, wz.x); }
#pragma unroll
                for (int u = 0; u < 4; ++u) dmma884(z0[u], z1[u], c[u].y, wz.y);
#pragma unroll
                for (int u = 0; u < 4; ++u) { x0[u] = 0.0; x1[u] = 0.0; dmma884(x0[u], x1[u], z0[u], wx.x); }
#pragma unroll
                for (int u = 0; u < 4; ++u) dmma884(x0[u], x1[u], z1[u], wx.y);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + u * nw;
                    double dot = fma(z1[u], zq.y, z0[u] * zq.x);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                    if (t < T) {
                        if (!DBG(2)) cp[9 * u * nw] = make_double2(x0[u], x1[u]);
                        if (!DBG(1)) { zp[2 * u * nw] = z0[u];
                        zp[2 * u * nw + VLEN] = z1[u]; }
                        if (m == 0 && !DBG(4)) rp[8 * u * nw] = (t == k) ? dot : rp[8 * u * nw] - dot;   // pivot tile: r_k <- W' zr = D^-1 r_k
                    }
                }
            }
        }
        PROF_TL(3);
        group_sync(gid, gthreads);
        PROF_TL(4);
        if (NMAT == 2 && gid == 0 && k == 0) asm volatile("bar.arrive 3, %0;" ::"r"(2 * gthreads) : "memory");

        // P4: A_ij -= Z_i Z_j' on every tile (tiles of row / column k are overwritten right after)
        if (!DBG(16))
#pragma unroll 1
        for (int kk = 0; kk < 8; ++kk) {
            const double* zb = Zb + kk * VLEN;
            double v[8];
            load8(zb, tj, v);
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const double2 x = *reinterpret_cast<const double2*>(zb + pp * CS + 2 * ti);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    A[2 * pp][c] = fma(-x.x, v[c], A[2 * pp][c]);
                    A[2 * pp + 1][c] = fma(-x.y, v[c], A[2 * pp + 1][c]);
                }
            }
        }
        PROF_TL(5);
        if (DBG(32)) {
        } else if (colo) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int slot = ti * 9 + r;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const double2 v = Cw[pp * PS + slot];
                    A[r][2 * pp] = v.x;
                    A[r][2 * pp + 1] = v.y;
                }
            }
        } else if (rowo) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int slot = tj * 9 + c;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const double2 v = Cw[pp * PS + slot];
                    A[2 * pp][c] = v.x;
                    A[2 * pp + 1][c] = v.y;
                }
            }
        } else if (dgo) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int slot = k * 9 + r;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const double2 v = Cw[pp * PS + slot];
                    A[r][2 * pp] = -v.x;
                    A[r][2 * pp + 1] = -v.y;
                }
            }
        }
        PROF_TL(6);
    }
    group_sync(gid, gthreads);
#ifdef GPCC_BLOCK_PROF
    if (blockIdx.x == 0) for (int i = tid; i < T * 8 * 12; i += gthreads) if ((i % 12) / 6 == gid) g_block_tl[i] = tlbuf[i];
#endif

    // ---- log-determinant, info, quadratic form -------------------------------------------------------------------
    double ld = 0.0;
    int bad = INT_MAX;
    for (int i = tid; i < N; i += gthreads) {
        const double d = piv[i];
        if (!(d > 0.0)) bad = min(bad, i + 1); else ld += log(d);
    }
    ld = group_sum(ld, red, tid, gthreads, gid);
    int* s_bad = reinterpret_cast<int*>(misc + 2);
    if (tid == 0) *s_bad = INT_MAX;
    group_sync(gid, gthreads);
    if (bad != INT_MAX) atomicMin(s_bad, bad);   // min is order independent: deterministic
    group_sync(gid, gthreads);
    const int info = (*s_bad == INT_MAX) ? 0 : *s_bad;
    const double quad = misc[0];
    const double ll = -0.5 * ((double)N * LOG2PI + ld + quad);   // logpdf(MvNormal(bbar,K), Y)  (:139)
    if (tid == 0) {
        b.ll[e] = info ? -INFINITY : ll;
        if (b.info) b.info[e] = info;
    }
    if (!b.want_grad) return;
    if (info) {
        if (tid <= L) b.grad[(size_t)e * (L + 1) + tid] = 0.0;
        return;
    }

    // ---- gradient: W = a a' - K~^-1 contracted with K and dK/drho ------------------------------------------------
    for (int i = tid; i < Np; i += gthreads) abuf[cidx(i)] = (i < N) ? rv[i] : 0.0;
    group_sync(gid, gthreads);   // also: nobody reads Zb / Cb any more, `part` may overwrite them

    if (b.dump_kinv && active) {   // K~^-1 = -(swept matrix); once per gpcc call (postb / pred), not in the fit loop
        double* out = b.dump_kinv + (size_t)e * N * N;
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int i = ti * 8 + r, j = tj * 8 + c;
                if (i < N && j < N && j <= i) { out[(size_t)j * N + i] = -A[r][c]; out[(size_t)i * N + j] = -A[r][c]; }
            }
    }
    if (b.dump_a) for (int i = tid; i < N; i += gthreads) b.dump_a[(size_t)e * N + i] = rv[i];

    double rows[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) rows[r] = 0.0;
    double es = 0.0;
    const bool diag_tile = (ti == tj);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        double tc[4], ac[4], wc[4], cols[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int cj = cidx(tj * 8 + half * 4 + cc);
            tc[cc] = tsh[cj]; ac[cc] = av[cj]; wc[cc] = abuf[cj]; cols[cc] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ci = cidx(ti * 8 + r);
            const double tr = tsh[ci], ar = av[ci], wr = abuf[ci];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = half * 4 + cc;
                const double Wv = fma(wr, wc[cc], A[r][c]);      // a_i a_j - (K~^-1)_ij
                double kv, dkv;
                kern_value_drho<KID>(tr - tc[cc], kp, kv, dkv);
                const double aa = ar * ac[cc];                   // 0 on padding rows
                double ct = Wv * (aa * kv);
                double et = Wv * (aa * dkv);
                if (diag_tile) {
                    if (r == c) { rows[r] += ct; ct = 0.0; et = 0.0; }   // diagonal counted once, dk(0)=0
                    else if (r < c) { ct = 0.0; et = 0.0; }              // upper part of the tile is unused
                }
                rows[r] += ct;
                cols[cc] += ct;
                es += et;
            }
        }
        if (active) {
            if (!diag_tile) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) part[(tj * T + ti) * 8 + half * 4 + cc] = cols[cc];
            } else {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) if (r == half * 4 + cc) rows[r] += cols[cc];
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < 8; ++r) part[(ti * T + tj) * 8 + r] = rows[r];
    }
    es = group_sum(active ? es : 0.0, red, tid, gthreads, gid);   // (its barriers also order `part`)

    // s_i = sum_j W_ij K_ij (full row);  dlogL/dalpha_p = (1/alpha_p) sum_{i in band p} s_i
    double* srow = rv;   // natural layout, reuse
    for (int i = tid; i < N; i += gthreads) {
        const double* pp = part + (size_t)(i >> 3) * T * 8 + (i & 7);
        double s = 0.0;
        for (int src = 0; src < T; ++src) s += pp[src * 8];
        srow[i] = s;
    }
    group_sync(gid, gthreads);
    const int warp = tid >> 5, lane = tid & 31, nwarps = gthreads >> 5;
    for (int pb = warp; pb < L; pb += nwarps) {
        double s = 0.0;
        for (int i = p.band_start[pb] + lane; i < p.band_start[pb + 1]; i += 32) s += srow[i];
        s = warp_sum(s);
        if (lane == 0) b.grad[(size_t)e * (L + 1) + pb] = s / b.alpha[(size_t)e * L + pb];
    }
    if (tid == 0) b.grad[(size_t)e * (L + 1) + L] = es;   // 0.5 * sum_full = sum over the strict lower triangle
}

size_t group_smem_doubles(int T, int want_grad) {
    const int Np = T * 8;
    const size_t zc = (size_t)8 * VLEN + (size_t)2 * 4 * (9 * T) * 2;
    const size_t part = want_grad ? (size_t)T * T * 8 : 0;
    size_t doubles = (size_t)VLEN * 7 + 128 + 8 + 8 + 16 + (zc > part ? zc : part) + (size_t)(Np + 1) / 2 + 2;
    return (doubles + 1) & ~(size_t)1;   // keep every group 16-byte aligned
}

template <int KID, int MAXTHREADS, int MINBLOCKS, int NMAT>
cudaError_t launch_variant(const DevProblem& p, const EvalBatch& b, int T, int gthreads, int Tmax, cudaStream_t s) {
    auto kfn = small_block_kernel<KID, MAXTHREADS, MINBLOCKS, NMAT>;
    const size_t gd = group_smem_doubles(T, b.want_grad);
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(group_smem_doubles(Tmax, 1) * 8 * NMAT) + 32 * 8 * 12 * 8);
    const int blocks = (b.M + NMAT - 1) / NMAT;
    size_t extra = 0;
#ifdef GPCC_BLOCK_PROF
    extra = 32 * 8 * 12 * 8;
#endif
    kfn<<<blocks, gthreads * NMAT, gd * 8 * NMAT + extra, s>>>(p, b, T, gthreads, (int)gd);
    return cudaGetLastError();
}

template <int KID>
cudaError_t launch_kid(const DevProblem& p, const EvalBatch& b, int T, cudaStream_t s) {
    const int ntiles = T * (T + 1) / 2, Np = 8 * T;
    const int gthreads = (ntiles + 31) / 32 * 32;
    (void)Np;
    static const int variant = getenv("GPCC_BLOCK_VARIANT") ? atoi(getenv("GPCC_BLOCK_VARIANT")) : 0;
    if (gthreads <= 128) return launch_variant<KID, 128, 3, 1>(p, b, T, gthreads, 15, s);
    if (gthreads <= 192) {
        if (variant == 1) return launch_variant<KID, 384, 1, 2>(p, b, T, gthreads, 19, s);
        return launch_variant<KID, 192, 2, 1>(p, b, T, gthreads, 19, s);
    }
    return launch_variant<KID, 352, 1, 1>(p, b, T, gthreads, MAX_T, s);
}

}  // namespace

bool small_block_supports(int N) { return (N + 7) / 8 <= MAX_T; }

cudaError_t small_block_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t s) {
    const int T = (p.N + 7) / 8;
#ifdef GPCC_BLOCK_PROF
    const int dbg = getenv("GPCC_BLOCK_DBG") ? atoi(getenv("GPCC_BLOCK_DBG")) : 0;
    cudaMemcpyToSymbol(g_block_dbg, &dbg, sizeof(dbg));
    cudaError_t rc = launch_kid<K_M32>(p, b, T, s);
    cudaStreamSynchronize(s);
    static long long tl[32 * 8 * 12];
    cudaMemcpyFromSymbol(tl, g_block_tl, sizeof(tl));
    const int nw = ((T * (T + 1) / 2 + 31) / 32);
    for (int g = 0; g < 2; ++g) {
        // per step: start = min slot0, B1 release ~ max slot1, B2 release ~ max slot3, bulk end = max slot5, step end = max slot6
        double acc[6] = {0}; long long first = 0, last = 0;
        for (int k = 1; k < T; ++k) {
            long long mn0 = 1LL << 62, mx1 = 0, mx3 = 0, mx5 = 0, mx6 = 0, mn5 = 1LL << 62, mn1 = 1LL << 62;
            for (int w = 0; w < nw; ++w) {
                const long long* q = tl + (k * 8) * 12 + w + 6 * g;
                mn0 = q[0] < mn0 ? q[0] : mn0; mx1 = q[12] > mx1 ? q[12] : mx1; mn1 = q[12] < mn1 ? q[12] : mn1; mx3 = q[36] > mx3 ? q[36] : mx3;
                mx5 = q[60] > mx5 ? q[60] : mx5; mn5 = q[60] < mn5 ? q[60] : mn5; mx6 = q[72] > mx6 ? q[72] : mx6;
            }
            if (k == 1) first = mn0;
            last = mx6;
            acc[0] += mx1 - mn0; acc[1] += mx3 - mx1; acc[2] += mx5 - mx3; acc[3] += mx6 - mx5; acc[4] += mn1 - mn0; acc[5] += mn5 - mx3;
        }
        if (last == 0) continue;
        fprintf(stderr, "[block tl] dbg=%d T=%d group %d per step: P1 %.0f (first warp done %.0f) P3 %.0f bulk %.0f (first warp %.0f) P4 %.0f | step %.0f\n", dbg, T, g,
                acc[0] / (T - 1), acc[4] / (T - 1), acc[1] / (T - 1), acc[2] / (T - 1), acc[5] / (T - 1), acc[3] / (T - 1), (double)(last - first) / (T - 1));
    }
    return rc;
#else
    switch (p.kernel_id) {
        case K_OU:  return launch_kid<K_OU>(p, b, T, s);
        case K_RBF: return launch_kid<K_RBF>(p, b, T, s);
        case K_M32: return launch_kid<K_M32>(p, b, T, s);
        case K_M52: return launch_kid<K_M52>(p, b, T, s);
    }
    return cudaErrorInvalidValue;
#endif
}

}  // namespace gpcc
