// Fused small-N evaluator, fragment form: the matrix lives in registers in DMMA accumulator-fragment layout, the
// symmetric sweep advances eight pivots per block step and every O(N^2)-per-step operation runs on the FP64 tensor pipe.
//
// Same reference code as small_sweep.cu: the objective
//   K = delayedCovariance(kernel, alpha, tau, rho, tarray) + Sobs + B ; logpdf(MvNormal(bbar, K), Y)
// (/root/reference/src/gpccfixdelay_marginaliseb.jl:133-141, src/delayedCovariance.jl:1-38) plus the analytic
// gradient 0.5 tr((a a' - K^-1) dK/dtheta).
//
// Why this form (measured on B200, profiles/README.md): a lone warp issues about one dependent instruction every four
// cycles, and at N = 150 only two matrices fit the register file of an SM, so whatever a block step does outside its
// trailing update runs at single-warp speed.  In the thread-per-tile layouts (small_sweep.cu, small_block.cu) the
// panel traffic is per THREAD (64 values through divergent branches) and costs more than the update itself.  Here
//   * the lower triangle of K~ (padded with identity pivots to a multiple of 8) is cut into 8x8 tiles dealt round-robin
//     to the warps of a group; a warp holds up to 32 tiles, lane (r = lane/4, m = lane%4) holding elements
//     (r, 2m) and (r, 2m+1) of each = the mma.sync.m8n8k4.f64 accumulator fragment.  Moving a tile to or from shared
//     memory is ONE 16-byte access per lane, warp-uniform; a tile row or column is spread evenly over the warps;
//   * block step k (pivots 8k..8k+7) of the sweep  [D C'; C R] -> [-D^-1, D^-1 C'; C D^-1, R - C D^-1 C']:
//       P3   per 8-row tile of the panel, on DMMA:  Z = C W'  (W = L^-1, D = L L', so Z Z' = C D^-1 C'),  X = Z W,
//            r -= Z zr.  The contraction index is permuted (k-halves take columns 2m, 2m+1) so that every operand
//            fragment is exactly what a lane holds or one 16-byte load: no shuffles, no transposes;
//       bulk A_ij -= Z_i Z_j' : two DMMA.8x8x4 per tile from two LDS.128, no barrier inside;
//       scan the tiles of row / column k take their final values X (pivot tile: -D^-1), the tiles of row / column k+1
//            are gathered into the next panel, and ONE lane factors the next pivot tile (Cholesky of the 8x8 block,
//            explicit inverse of the TRIANGULAR factor only) while the other warps are still in their update;
//     two barriers per eight pivots;
//   * several matrices per CTA (named barriers per matrix, staggered start), so that the serial part of one matrix's
//     block step runs under another's DMMA stream, with the warps of the matrices interleaved over the four schedulers.
//   Work: N^3 flop per logL+grad evaluation as before (+ ~15 % for P3).
#include "gpcc_internal.h"
#include "kernfun.cuh"
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace gpcc {

#ifdef GPCC_FRAG_PROF
// Dev build only: per-warp timestamps of every block step, kept in shared memory and dumped at exit.
__device__ long long g_frag_tl[32 * 8 * 12];
#define PROF_TL(slot) do { if (lane == 0) tlbuf[(k * 8 + (slot)) * 12 + warp + nwarps * gid] = clock64(); } while (0)
#else
#define PROF_TL(slot)
#endif

namespace {

constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr int MAX_T = 25;      // 8 * 25 = 200 rows (N <= 199 as in small_sweep.cu); 325 tiles = 11 warps
constexpr int MAX_THREADS = 384;
constexpr int P3_TILES = 2;   // panel tiles in flight per warp in P3 (registers: the 128 accumulators leave ~40)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// Shared memory of one matrix ("group"), byte offsets from the group base.  Everything the block sweep touches sits at
// a compile-time offset (plus multiples of the panel size pb = 512 T), and is addressed through 32-bit shared-window
// addresses with explicit ld/st.shared: with 128 accumulator registers live the compiler otherwise rebuilds every
// generic pointer from scratch (S2R, S2UR, a dozen dependent IMADs) in front of each access, which is what made every
// phase of the first version run at ~10 cycles per instruction.
constexpr int SLOTS = 32;       // tiles per warp ...
constexpr int REG_SLOTS = 28;   // ... of which this many live in registers (112 of the 168 a thread may use at 12 warps
                                // per SM); the other four stay in shared memory and take one LDS.128 + STS.128 per block
                                // step.  With all 32 in registers ptxas keeps four of them in LOCAL memory instead.
constexpr unsigned O_W = 0, O_WT = 512, O_ZR = 1024, O_MISC = 1088, O_RED = 1152, O_DB = 1280, O_DN = 1792, O_RV = 2304;
constexpr unsigned O_PIV = O_RV + 1600, O_PARK = O_PIV + 1600, O_TM = O_PARK + REG_SLOTS * 512;
// O_TM: [nwarps][SLOTS - REG_SLOTS] memory-resident tiles; then Zb, Zn, Cb[0], Cb[1] (pb bytes each)
__device__ __forceinline__ double2 lds128(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts64(unsigned a, double x) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory");
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
__device__ __forceinline__ void tile_of(int q, int& ti, int& tj) {
    ti = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
    while (ti * (ti + 1) / 2 > q) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
    tj = q - ti * (ti + 1) / 2;
}

// Reciprocal square root of a positive double: hardware seed and two Newton steps (~1.5 ulp); a non-positive pivot gives
// NaN, which is what flags the matrix as not positive definite.
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

// Pivot tile of block step k, one lane: D (row major, full) -> L (D = L L'), W = L^-1, the Schur pivots (log-det and
// LAPACK-style info), zr = W r_k and the running quadratic form r' K~^-1 r.  `gb` = shared address of the group.
__device__ __forceinline__ void pivot_tile(unsigned gb, int k) {
    double A[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) A[i][j] = lds64(gb + O_DB + (i * 8 + j) * 8);
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = lds64(gb + O_RV + k * 64 + q * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double d = A[j][j];
        sts64(gb + O_PIV + k * 64 + j * 8, d);
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
    double qs = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double s = A[c][c] * r[c];
#pragma unroll
        for (int q = 0; q < c; ++q) s = fma(A[q][c], r[q], s);
        sts64(gb + O_ZR + c * 8, s);
        qs = fma(s, s, qs);
#pragma unroll
        for (int q = 0; q <= c; ++q) {
            const double w = (q == c) ? A[c][c] : A[q][c];
            sts64(gb + O_W + (c * 8 + q) * 8, w);     // W  row major (zeros above the diagonal, written once at start)
            sts64(gb + O_WT + (q * 8 + c) * 8, w);    // W' row major
        }
    }
    sts64(gb + O_MISC, lds64(gb + O_MISC) + qs);
}

// Tile s of the warp: registers for s < REG_SLOTS, shared memory (tm = this lane's pair of the warp's first memory tile)
#define TILE_GET(s, v0, v1) do { if ((s) < REG_SLOTS) { v0 = acc[(s) < REG_SLOTS ? (s) : 0][0]; v1 = acc[(s) < REG_SLOTS ? (s) : 0][1]; } \
                                 else { const double2 t_ = lds128(tm + ((s) - REG_SLOTS) * 512); v0 = t_.x; v1 = t_.y; } } while (0)
#define TILE_PUT(s, v0, v1) do { if ((s) < REG_SLOTS) { acc[(s) < REG_SLOTS ? (s) : 0][0] = v0; acc[(s) < REG_SLOTS ? (s) : 0][1] = v1; } \
                                 else sts128(tm + ((s) - REG_SLOTS) * 512, v0, v1); } while (0)

// One pass over the tiles of a warp after the trailing update of block step k (kw = k), or before the first step
// (kw = -1).  Warp-uniform tests only; a tile costs one or two shared-memory accesses per lane:
//   * tiles of row / column kw take their final values X = C D^-1 from the panel buffer (pivot tile: -D^-1);
//   * tiles of row / column kw+1 are copied into the next panel buffer; its pivot tile leaves identity rows there
//     (their X rows become D^-1); when `first`, the pivot tile itself goes to Db (later ones arrive through P3);
//   * the pivot tile kw+2 is copied, as it stands, to Dn: P3 of step kw+1 finishes it (look-ahead).
__device__ __forceinline__ void scan_tiles(double (&acc)[REG_SLOTS][2], unsigned tm, int mytile, int kw, unsigned lane16, unsigned lanet,
                                           unsigned cw, unsigned cn, unsigned gb, bool first) {
    // lane16 = 16 * lane: this lane's pair inside a row-major tile; lanet = 8 * (lc * 8 + lr): the same pair transposed
    const int mti = mytile >> 8, mtj = mytile & 255;
    const int kn = kw + 1, kp = kw + 2;
    const unsigned mask = __ballot_sync(0xffffffffu, mytile >= 0 && (mti == kw || mtj == kw || mti == kn || mtj == kn || (mti == kp && mtj == kp)));
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        if ((mask >> s) & 1u) {
            const int tt = __shfl_sync(0xffffffffu, mytile, s);
            const int ti = tt >> 8, tj = tt & 255;
            double v0, v1;
            if (tj == kw) {           // column tile: X rows ti*8 + r;  pivot tile: -D^-1
                const double2 v = lds128(cw + ti * 512 + lane16);
                const double sg = (ti == kw) ? -1.0 : 1.0;
                v0 = sg * v.x;
                v1 = sg * v.y;
                TILE_PUT(s, v0, v1);
            } else if (ti == kw) {    // row tile (kw, tj): element (r, c) = X[tj*8 + c][r]
                v0 = lds64(cw + tj * 512 + lanet);
                v1 = lds64(cw + tj * 512 + lanet + 64);
                TILE_PUT(s, v0, v1);
            } else {
                TILE_GET(s, v0, v1);
            }
            if (tj == kn) {
                if (ti == kn) {       // next pivot tile: identity rows in the panel
                    const int lr = lane16 >> 6, lc = (lane16 >> 3) & 6;
                    sts128(cn + kn * 512 + lane16, lr == lc ? 1.0 : 0.0, lr == lc + 1 ? 1.0 : 0.0);
                    if (first) sts128(gb + O_DB + lane16, v0, v1);
                } else {              // column tile: panel rows ti*8 + r
                    sts128(cn + ti * 512 + lane16, v0, v1);
                }
            } else if (ti == kn) {    // row tile (kn, tj): panel rows tj*8 + c, entry r
                sts64(cn + tj * 512 + lanet, v0);
                sts64(cn + tj * 512 + lanet + 64, v1);
            } else if (ti == kp && tj == kp) {
                sts128(gb + O_DN + lane16, v0, v1);
            }
        }
    }
}

template <int KID>
__global__ void __launch_bounds__(MAX_THREADS, 1)
small_frag_kernel(DevProblem p, EvalBatch b, int T, int gthreads, int nmat, int smem_doubles_per_group) {
    extern __shared__ __align__(16) double smem_all[];
    const int N = p.N, L = p.L;
    const int Np = 8 * T;
    const int gid = (int)threadIdx.x / gthreads;
    const int tid = (int)threadIdx.x - gid * gthreads;
    const int e = blockIdx.x * nmat + gid;
    if (e >= b.M) return;   // whole group leaves: its named barriers are never used
    const bool next_group_exists = (gid + 1 < nmat) && (e + 1 < b.M);
    double* smem = smem_all + (size_t)gid * smem_doubles_per_group;
    const int lane = tid & 31, warp = tid >> 5, nwarps = gthreads >> 5;
    const int ntiles = T * (T + 1) / 2;
    const int lr = lane >> 2, lc = (lane & 3) * 2;   // this lane's row and first column inside a tile

    const unsigned pb = 512u * T;                                   // bytes of one panel buffer
    const unsigned gb = (unsigned)__cvta_generic_to_shared(smem);   // shared-window address of this group
    const unsigned lane16 = 16u * lane, lanet = 8u * (lc * 8 + lr);
    double* Wb = smem + O_W / 8;        // [64] W = L^-1 row major, [64] W' row major
    double* misc = smem + O_MISC / 8;   // [0] quadratic form, [2] info (as int)
    double* red = smem + O_RED / 8;     // [16] warp totals
    double* rv = smem + O_RV / 8;       // [Np] residual, swept along with the matrix: ends as a = K~^-1 r
    double* piv = smem + O_PIV / 8;     // [Np] Schur pivots
    const unsigned ozb = O_TM + (unsigned)nwarps * (SLOTS - REG_SLOTS) * 512;
    const unsigned tm = gb + O_TM + (unsigned)warp * (SLOTS - REG_SLOTS) * 512 + lane16;   // this lane's pair of the warp's first memory tile
    double* Zb = smem + ozb / 8;        // [T][64] Z, then [T][64] -Z, then [2][T][64] panel C / X = C D^-1 (row-major 8x8 per tile row)
    double2* stage = reinterpret_cast<double2*>(Zb);   // [nwarps][8][32] fragment staging (assembly, gradient): aliases the panels
    const size_t panel_doubles = (size_t)4 * T * 64, stage_doubles = (size_t)nwarps * 512;
    double* tsh = Zb + (panel_doubles > stage_doubles ? panel_doubles : stage_doubles);   // [Np] shifted times
    double* av = tsh + Np;              // [Np] alpha per point (0 on padding)
    double* sbv = av + Np;              // [Np] Sigma_b per point
    double* dadd = sbv + Np;            // [Np] sigma^2 (1 on padding: identity pivots)
    double* part = dadd + Np;           // [T][T][8] gradient partials
    double* partd = part + (b.want_grad ? T * T * 8 : 0);                                  // [T][8]
    int* bandv = reinterpret_cast<int*>(partd + (b.want_grad ? T * 8 : 0));                // [Np]
#ifdef GPCC_FRAG_PROF
    long long* tlbuf = reinterpret_cast<long long*>(smem_all + (size_t)nmat * smem_doubles_per_group);
#endif
    double2* st = stage + warp * 256;

    const double rho = b.rho[e];
    const KernParams kp = make_kern_params(KID, rho);

    for (int i = tid; i < Np; i += gthreads) {
        double ts = 0.0, al = 0.0, sb = 0.0, dd = 1.0, r = 0.0;
        int bi = -1 - i;
        if (i < N) {
            bi = p.band[i];
            ts = p.t[i] - b.delays[(size_t)e * L + bi];        // delayedCovariance.jl:27 (x - delays[l])
            al = b.alpha[(size_t)e * L + bi];
            sb = b.mode_postb ? 0.0 : p.sigb[i];
            dd = p.s2[i];
            r = b.mode_postb ? p.y[i] : p.resid[i];
        }
        tsh[i] = ts; av[i] = al; sbv[i] = sb; dadd[i] = dd; rv[i] = r; bandv[i] = bi;
    }
    if (tid < 128) Wb[tid] = 0.0;
    if (tid == 0) misc[0] = 0.0;
    // slot s of this warp is tile q = s * nwarps + warp (round robin); lane s remembers it for the whole warp
    int mytile = -1;
    {
        const int q = lane * nwarps + warp;
        if (q < ntiles) { int a_, b_; tile_of(q, a_, b_); mytile = (a_ << 8) | b_; }
    }
    const unsigned myoffs = mytile < 0 ? 0u : (unsigned)((mytile >> 8) * 512) | ((unsigned)((mytile & 255) * 512) << 16);
    group_sync(gid, gthreads);

    // ---- assembly, eight tiles at a time through the staging buffer (keeps the exp code out of the unrolled part) ---
    double acc[REG_SLOTS][2];
#pragma unroll
    for (int c = 0; c < SLOTS / 8; ++c) {
        for (int u = 0; u < 8; ++u) {
            const int q = (c * 8 + u) * nwarps + warp;
            double v0 = 0.0, v1 = 0.0;
            if (q < ntiles) {
                int ti, tj;
                tile_of(q, ti, tj);
                const int i = ti * 8 + lr, j = tj * 8 + lc;
                const double ai = av[i], ti_ = tsh[i];
                v0 = (ai * av[j]) * kern_value<KID>(ti_ - tsh[j], kp);          // scale[l]*scale[m]*kernel (delayedCovariance.jl:27)
                v1 = (ai * av[j + 1]) * kern_value<KID>(ti_ - tsh[j + 1], kp);
                if (i == j) v0 += dadd[i];                                       // + Sobs  (gpccfixdelay_marginaliseb.jl:135)
                if (i == j + 1) v1 += dadd[i];
                if (bandv[i] == bandv[j]) v0 += sbv[i];                          // + B = Q Sigma_b Q'
                if (bandv[i] == bandv[j + 1]) v1 += sbv[i];
            }
            st[u * 32 + lane] = make_double2(v0, v1);
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double2 v = st[u * 32 + lane];
            TILE_PUT(c * 8 + u, v.x, v.y);
        }
        __syncwarp();
    }
    group_sync(gid, gthreads);   // the staging buffer aliases the panel buffers

    // ---- block sweep ----------------------------------------------------------------------------------------------
    // Matrices of one CTA start one after the other: the next one enters its sweep when this one reaches its first
    // trailing update, so that their serial phases and DMMA streams interleave from then on.
    if (gid > 0) asm volatile("bar.sync %0, %1;" ::"r"(8 + gid), "r"(2 * gthreads) : "memory");
    scan_tiles(acc, tm, mytile, -1, lane16, lanet, gb + ozb + 2 * pb, gb + ozb + 2 * pb, gb, true);   // panel 0, pivot tile 0 -> Db, pivot tile 1 -> Dn
    if (warp == 0) {   // tile (0,0) is slot 0 of warp 0
        __syncwarp();
        if (lane == 0) pivot_tile(gb, 0);
    }
    int pw = (nwarps > 1) ? 1 : 0;   // warp that finishes pivot tile k+1 in P3 of step k: (k+1) % nwarps
    for (int k = 0; k < T; ++k) {
        const unsigned cw = gb + ozb + (2 + (k & 1)) * pb, cn = gb + ozb + (3 - (k & 1)) * pb;
        PROF_TL(0);
        group_sync(gid, gthreads);
        PROF_TL(1);
        // P3: Z = C W', X = Z W, r -= Z zr on DMMA, up to four independent 8-row tiles in flight per warp
        {
            const double2 wz = lds128(gb + O_W + lane16);     // W[n=lr][2m, 2m+1]
            const double2 wx = lds128(gb + O_WT + lane16);    // W[2m, 2m+1][n=lr]
            const double2 zq = lds128(gb + O_ZR + (lane16 & 48));
            for (int t0 = warp; t0 < T; t0 += P3_TILES * nwarps) {
                const unsigned ca = cw + t0 * 512 + lane16, tstep = 512u * nwarps;
                double2 c[P3_TILES];
                double z0[P3_TILES], z1[P3_TILES], x0[P3_TILES], x1[P3_TILES], rold[P3_TILES];
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) c[u] = (t0 + u * nwarps < T) ? lds128(ca + u * tstep) : make_double2(0.0, 0.0);
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) rold[u] = (t0 + u * nwarps < T) ? lds64(gb + O_RV + (t0 + u * nwarps) * 64 + lr * 8) : 0.0;
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) { z0[u] = 0.0; z1[u] = 0.0; dmma884(z0[u], z1[u], c[u].x, wz.x); }
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) dmma884(z0[u], z1[u], c[u].y, wz.y);
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) { x0[u] = 0.0; x1[u] = 0.0; dmma884(x0[u], x1[u], z0[u], wx.x); }
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) dmma884(x0[u], x1[u], z1[u], wx.y);
                double dot[P3_TILES];
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) dot[u] = fma(z1[u], zq.y, z0[u] * zq.x);
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], 1);
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], 2);
#pragma unroll
                for (int u = 0; u < P3_TILES; ++u) {
                    const int t = t0 + u * nwarps;
                    if (t < T) {
                        sts128(ca + u * tstep, x0[u], x1[u]);
                        sts128(ca + u * tstep - (2 + (k & 1)) * pb, z0[u], z1[u]);          // Zb
                        sts128(ca + u * tstep - (1 + (k & 1)) * pb, -z0[u], -z1[u]);        // Zn
                        if ((lane & 3) == 0) sts64(gb + O_RV + t * 64 + lr * 8, (t == k) ? dot[u] : rold[u] - dot[u]);   // pivot tile: r_k <- W' zr = D^-1 r_k
                        if (t == k + 1) {   // look-ahead: the next pivot tile after this step's update, D - Z Z'
                            double2 d = lds128(gb + O_DN + lane16);
                            dmma884(d.x, d.y, -z0[u], z0[u]);
                            dmma884(d.x, d.y, -z1[u], z1[u]);
                            sts128(gb + O_DB + lane16, d.x, d.y);
                        }
                    }
                }
            }
        }
        PROF_TL(2);
        group_sync(gid, gthreads);
        PROF_TL(3);
        if (k == 0 && next_group_exists) asm volatile("bar.arrive %0, %1;" ::"r"(8 + gid + 1), "r"(2 * gthreads) : "memory");

        // look-ahead: ONE lane of the warp that finished the next pivot tile factors it (Cholesky of the 8x8 block,
        // W = L^-1, zr) while the other warps run their trailing update.  Its tile registers are parked in shared
        // memory meanwhile: the factorisation needs the registers of a whole thread.
        if (k + 1 < T && warp == pw) {
#pragma unroll
            for (int s = 0; s < REG_SLOTS; ++s) sts128(gb + O_PARK + s * 512 + lane16, acc[s][0], acc[s][1]);
            if (lane == 0) pivot_tile(gb, k + 1);
#pragma unroll
            for (int s = 0; s < REG_SLOTS; ++s) { const double2 v = lds128(gb + O_PARK + s * 512 + lane16); acc[s][0] = v.x; acc[s][1] = v.y; }
        }
        pw = (pw + 1 == nwarps) ? 0 : pw + 1;
        PROF_TL(4);
        // bulk: A_ij -= Z_i Z_j' on every tile of the warp (tiles of row / column k are overwritten right after);
        // operands are fetched one tile ahead (a warp issues in order: LDS and SHFL latencies must be covered by hand)
        {
            const unsigned zbb = gb + ozb + lane16, znb = zbb + pb;
            unsigned o0 = __shfl_sync(0xffffffffu, myoffs, 0);
            double2 a_0 = lds128(znb + (o0 & 0xffffu)), b_0 = lds128(zbb + (o0 >> 16));
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                double2 a_1 = a_0, b_1 = b_0;
                if (s + 1 < SLOTS) {
                    const unsigned o1 = __shfl_sync(0xffffffffu, myoffs, s + 1);
                    a_1 = lds128(znb + (o1 & 0xffffu));
                    b_1 = lds128(zbb + (o1 >> 16));
                }
                if (s < REG_SLOTS) {
                    dmma884(acc[s < REG_SLOTS ? s : 0][0], acc[s < REG_SLOTS ? s : 0][1], a_0.x, b_0.x);
                    dmma884(acc[s < REG_SLOTS ? s : 0][0], acc[s < REG_SLOTS ? s : 0][1], a_0.y, b_0.y);
                } else {
                    double2 d = lds128(tm + (s - REG_SLOTS) * 512);
                    dmma884(d.x, d.y, a_0.x, b_0.x);
                    dmma884(d.x, d.y, a_0.y, b_0.y);
                    sts128(tm + (s - REG_SLOTS) * 512, d.x, d.y);
                }
                a_0 = a_1; b_0 = b_1;
            }
        }
        PROF_TL(5);
        // scan: final values of row / column k, next panel, raw pivot tile k+2
        scan_tiles(acc, tm, mytile, k, lane16, lanet, cw, cn, gb, false);
        PROF_TL(6);
    }
    group_sync(gid, gthreads);
#ifdef GPCC_FRAG_PROF
    if (blockIdx.x == 0) for (int i = tid; i < T * 8 * 12; i += gthreads) if ((i % 12) / nwarps == gid) g_frag_tl[i] = tlbuf[i];
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
    if (b.dump_a) for (int i = tid; i < N; i += gthreads) b.dump_a[(size_t)e * N + i] = rv[i];

    // ---- gradient: W = a a' - K~^-1 contracted with K and dK/drho, eight tiles at a time through the staging buffer -
    double es = 0.0;
#pragma unroll
    for (int c = 0; c < SLOTS / 8; ++c) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { double g0, g1; TILE_GET(c * 8 + u, g0, g1); st[u * 32 + lane] = make_double2(g0, g1); }
        __syncwarp();
        for (int u = 0; u < 8; ++u) {
            const int q = (c * 8 + u) * nwarps + warp;
            if (q < ntiles) {
                int ti, tj;
                tile_of(q, ti, tj);
                const double2 ainv = st[u * 32 + lane];              // -(K~^-1) entries of this lane
                const double am[2] = {ainv.x, ainv.y};
                const int i = ti * 8 + lr;
                double ct[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = tj * 8 + lc + h;
                    const double Wv = fma(rv[i], rv[j], am[h]);       // a_i a_j - (K~^-1)_ij
                    double kv, dkv;
                    kern_value_drho<KID>(tsh[i] - tsh[j], kp, kv, dkv);
                    const double aa = av[i] * av[j];                  // 0 on padding rows
                    double cc = Wv * (aa * kv), d = Wv * (aa * dkv);
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
        }
        __syncwarp();
    }
    es = group_sum(es, red, tid, gthreads, gid);     // its barriers also order `part`
    // s_i = sum_j W_ij K_ij (full row);  dlogL/dalpha_p = (1/alpha_p) sum_{i in band p} s_i
    double* srow = piv;                              // reuse
    for (int i = tid; i < N; i += gthreads) {
        const int ti = i >> 3, r = i & 7;
        double s = partd[ti * 8 + r];
        for (int src = 0; src < T; ++src) s += part[(ti * T + src) * 8 + r];
        srow[i] = s;
    }
    group_sync(gid, gthreads);
    for (int pb = warp; pb < L; pb += nwarps) {
        double s = 0.0;
        for (int i = p.band_start[pb] + lane; i < p.band_start[pb + 1]; i += 32) s += srow[i];
        s = warp_sum(s);
        if (lane == 0) b.grad[(size_t)e * (L + 1) + pb] = s / b.alpha[(size_t)e * L + pb];
    }
    if (tid == 0) b.grad[(size_t)e * (L + 1) + L] = es;   // 0.5 * sum_full = sum over the strict lower triangle
}

size_t group_smem_doubles(int T, int nwarps, int want_grad) {
    const int Np = 8 * T;
    const size_t panel = (size_t)4 * T * 64, stage = (size_t)nwarps * 512;
    size_t doubles = (O_TM + (size_t)nwarps * (SLOTS - REG_SLOTS) * 512) / 8 + (panel > stage ? panel : stage) + (size_t)Np * 4 +
                     (want_grad ? (size_t)T * T * 8 + (size_t)T * 8 : 0) + (size_t)(Np + 1) / 2 + 2;
    return (doubles + 1) & ~(size_t)1;   // keep every group 16-byte aligned
}

template <int KID>
cudaError_t launch_kid(const DevProblem& p, const EvalBatch& b, int T, cudaStream_t s) {
    const int ntiles = T * (T + 1) / 2;
    const int nwarps = (ntiles + SLOTS - 1) / SLOTS;
    const int gthreads = nwarps * 32;
    static const int nmat_cap = getenv("GPCC_FRAG_NMAT") ? atoi(getenv("GPCC_FRAG_NMAT")) : 4;
    int nmat = MAX_THREADS / gthreads;
    if (nmat > nmat_cap) nmat = nmat_cap;
    if (nmat < 1) nmat = 1;
    const size_t gd = group_smem_doubles(T, nwarps, b.want_grad);
    auto kfn = small_frag_kernel<KID>;
    size_t extra = 0;
#ifdef GPCC_FRAG_PROF
    extra = 32 * 8 * 12 * 8;
#endif
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_done = true; }
    const int blocks = (b.M + nmat - 1) / nmat;
    kfn<<<blocks, gthreads * nmat, gd * 8 * nmat + extra, s>>>(p, b, T, gthreads, nmat, (int)gd);
    return cudaGetLastError();
}

}  // namespace

bool small_frag_supports(int N) { return (N + 7) / 8 <= MAX_T; }

cudaError_t small_frag_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t s) {
    const int T = (p.N + 7) / 8;
#ifdef GPCC_FRAG_PROF
    cudaError_t rc = launch_kid<K_M32>(p, b, T, s);
    cudaStreamSynchronize(s);
    static long long tl[32 * 8 * 12];
    cudaMemcpyFromSymbol(tl, g_frag_tl, sizeof(tl));
    const int nw = ((T * (T + 1) / 2 + 31) / 32);
    for (int g = 0; g < 12 / nw && g < 4; ++g) {
        // slots: 0 step start, 1 after B1, 2 P3 done, 3 after B2, 4 look-ahead pivot done, 5 bulk done, 6 scan done
        double acc[6] = {0}; long long first = 0, last = 0;
        for (int k = 1; k < T; ++k) {
            long long mn0 = 1LL << 62, mx[8] = {0};
            for (int w = 0; w < nw; ++w) {
                const long long* q = tl + (k * 8) * 12 + w + nw * g;
                mn0 = q[0] < mn0 ? q[0] : mn0;
                for (int sl = 0; sl < 7; ++sl) mx[sl] = q[12 * sl] > mx[sl] ? q[12 * sl] : mx[sl];
            }
            if (k == 1) first = mn0;
            last = mx[6];
            acc[0] += mx[1] - mn0; acc[1] += mx[3] - mx[1]; acc[2] += mx[4] - mx[3]; acc[3] += mx[5] - mx[4]; acc[4] += mx[6] - mx[5];
        }
        if (last == 0) continue;
        if (g == 0) for (int w = 0; w < nw; ++w) {
            const long long* q = tl + (5 * 8) * 12 + w;
            fprintf(stderr, "[frag tl] T=%d k=5 warp %d: start %lld | after B1 %lld | P3 done %lld | after B2 %lld | pivot done %lld | bulk done %lld | scan done %lld\n", T, w,
                    q[0] - tl[5 * 8 * 12], q[12] - tl[5 * 8 * 12], q[24] - tl[5 * 8 * 12], q[36] - tl[5 * 8 * 12], q[48] - tl[5 * 8 * 12], q[60] - tl[5 * 8 * 12], q[72] - tl[5 * 8 * 12]);
        }
        fprintf(stderr, "[frag tl] T=%d group %d per step (slowest warp): B1 %.0f | P3+B2 %.0f | pivot %.0f | bulk %.0f | scan %.0f | step %.0f\n", T, g,
                acc[0] / (T - 1), acc[1] / (T - 1), acc[2] / (T - 1), acc[3] / (T - 1), acc[4] / (T - 1), (double)(last - first) / (T - 1));
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
