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
__device__ long long g_frag_tl[32 * 8 * 16];
#define PROF_TL(slot) do { if (lane == 0) tlbuf[(k * 8 + (slot)) * 16 + warp + (nwarps + 1) * gid] = clock64(); } while (0)
#else
#define PROF_TL(slot)
#endif

namespace {

constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr int MAX_T = 25;      // 8 * 25 = 200 rows (N <= 199 as in small_sweep.cu); 325 tiles = 11 warps
constexpr int MAX_THREADS = 512;   // HELPER build: 16 warps x 128 registers (registers are handed out four warps at a time)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// Shared memory of one matrix ("group"), byte offsets from the group base.  Everything the block sweep touches sits at
// a compile-time offset (plus multiples of the panel size pb = 512 T), and is addressed through 32-bit shared-window
// addresses with explicit ld/st.shared: with 128 accumulator registers live the compiler otherwise rebuilds every
// generic pointer from scratch (S2R, S2UR, a dozen dependent IMADs) in front of each access, which is what made every
// phase of the first version run at ~10 cycles per instruction.
constexpr int SLOTS = 32;       // tiles per warp ...
// ... of which RS live in registers; the others stay in shared memory and take one LDS.128 + STS.128 per block step (with all
// 32 in registers ptxas keeps some of them in LOCAL memory instead, which is worse).  Two builds of the kernel:
//   HELPER = false: 6 tile warps per matrix at N = 150, one matrix per CTA, two CTAs per SM at 168 registers, RS = 28; the
//                   look-ahead pivot factorisation runs on the tile warp that finished the pivot tile (its tiles parked);
//   HELPER = true:  a seventh warp per matrix does the factorisation on a scheduler of its own; registers are granted four
//                   warps at a time, so this means 16 warps x 128 registers per SM and RS = 20.
template <bool HELPER> struct FragCfg { static constexpr int RS = HELPER ? 20 : 28; };
constexpr int CHUNK = 4;        // consecutive tiles (row major over the lower triangle) dealt to a warp at a time: the
                                // A operand of the update is reloaded only when the tile row changes
constexpr unsigned O_W = 0, O_WT = 512, O_ZR = 1024, O_MISC = 1088, O_RED = 1152, O_DB = 1280, O_DN = 1792, O_RV = 2816;   // Dn: two buffers
constexpr unsigned O_PIV = O_RV + 3072, O_TM = O_PIV + 2048;
// from O_TM on: [nwarps][SLOTS - RS] memory tiles, (HELPER = false) [RS] parked tiles, descriptors, operand table, panels
// from O_TM on (sizes depend on the launch): [nwarps][4] memory-resident tiles, [nwarps][32] scan descriptors (8 B),
// [nwarps][32] update-operand offsets (8 B), then Zb, Zn, Cb[0], Cb[1] (pb = 512 Tpad bytes each)
__device__ __forceinline__ uint2 lds64u(unsigned a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64u(unsigned a, unsigned x, unsigned y) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ double negd(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }   // ALU, not FP64 pipe
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
    }
    // W row major and W' row major, in 16-byte pieces (W[c][q] for q < c is A[q][c], W[c][c] = A[c][c], zero above)
#define W_AT(c, q) ((q) > (c) ? 0.0 : ((q) == (c) ? A[c][c] : A[q][c]))
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int q = 0; q <= c; q += 2) sts128(gb + O_W + (c * 8 + q) * 8, W_AT(c, q), W_AT(c, q + 1));
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = q & ~1; c < 8; c += 2) sts128(gb + O_WT + (q * 8 + c) * 8, W_AT(c, q), W_AT(c + 1, q));
#undef W_AT
    sts64(gb + O_MISC, lds64(gb + O_MISC) + qs);
}

// Tile s of the warp: registers for s < RS, shared memory (tm = this lane's pair of the warp's first memory tile)
#define TILE_GET(s, v0, v1) do { if ((s) < RS) { v0 = acc[(s) < RS ? (s) : 0][0]; v1 = acc[(s) < RS ? (s) : 0][1]; } \
                                 else { const double2 t_ = lds128(tm + ((s) - RS) * 512); v0 = t_.x; v1 = t_.y; } } while (0)
#define TILE_PUT(s, v0, v1) do { if ((s) < RS) { acc[(s) < RS ? (s) : 0][0] = v0; acc[(s) < RS ? (s) : 0][1] = v1; } \
                                 else sts128(tm + ((s) - RS) * 512, v0, v1); } while (0)

// Scan descriptors of block step kw (kw = -1 before the first step), one per tile, written by the lane whose number is
// the tile's slot: {address to take the final value from, address to copy the tile to} (0 = nothing to do).
//   * tiles of column kw (and the pivot tile) read X = C D^-1 (pivot: -D^-1) from block ti of the panel buffer, tiles of
//     row kw from block tj: P3 stores every block in the fragment layout of the tile that will read it;
//   * tiles of column / row kw+1 are copied to block ti / tj of the next panel buffer, as they are (P3 reads the blocks
//     that come from row tiles transposed);  the pivot tile kw+2 goes to Dn (look-ahead), and before the first step the
//     pivot tile 0 goes to Db.
__device__ __forceinline__ void write_descriptors(unsigned da, int mytile, int kw, unsigned cw, unsigned cn, unsigned gb) {
    const int mti = mytile >> 8, mtj = mytile & 255, kn = kw + 1, kp = kw + 2;
    unsigned src = 0, dst = 0;
    if (mytile >= 0) {
        if (mtj == kw) src = cw + mti * 512; else if (mti == kw) src = cw + mtj * 512;
        if (mtj == kn) { if (mti != kn) dst = cn + mti * 512; else if (kw < 0) dst = gb + O_DB; }
        else if (mti == kn) dst = cn + mtj * 512;
        else if (mti == kp && mtj == kp) dst = gb + O_DN + (kw & 1) * 512;   // read by the helper during step kw+1, while the scan of that step fills the other buffer
    }
    sts64u(da, src, dst);
}
// The scan proper: one predicated 16-byte load and one predicated 16-byte store per tile, no branches for the tiles
// in registers.
template <int RS>
__device__ __forceinline__ void scan_tiles(double (&acc)[RS][2], unsigned tm, unsigned dw, unsigned lane16) {
    // A warp issues in order and these accesses are ordered among themselves: descriptors are fetched eight at a time,
    // then the loads they select, then the stores, so that the latencies overlap instead of adding up.
#pragma unroll
    for (int g = 0; g < SLOTS; g += 8) {
        uint2 d[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = lds64u(dw + (g + u) * 8);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int s = g + u;
            if (s < RS) {
                if (d[u].x) { const double2 v = lds128(d[u].x + lane16); acc[s < RS ? s : 0][0] = v.x; acc[s < RS ? s : 0][1] = v.y; }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int s = g + u;
            if (s < RS) {
                if (d[u].y) sts128(d[u].y + lane16, acc[s < RS ? s : 0][0], acc[s < RS ? s : 0][1]);
            } else if (d[u].x | d[u].y) {
                const double2 v = lds128(d[u].x ? d[u].x + lane16 : tm + (s - RS) * 512);
                if (d[u].x) sts128(tm + (s - RS) * 512, v.x, v.y);
                if (d[u].y) sts128(d[u].y + lane16, v.x, v.y);
            }
        }
    }
}

template <int KID, bool HELPER>
__global__ void __launch_bounds__(HELPER ? MAX_THREADS : 384, 1)
small_frag_kernel(DevProblem p, EvalBatch b, int T, int Tpad, int gthreads, int hw_per_group, int nmat, int smem_doubles_per_group) {
    constexpr int RS = FragCfg<HELPER>::RS;
    extern __shared__ __align__(16) double smem_all[];
    const int N = p.N, L = p.L;
    const int Np = 8 * T;
    // Hardware warp w runs on scheduler w % 4.  Tile warps take the slots with w % 4 != 3, the helper warp of each matrix
    // one with w % 4 == 3: its scalar FP64 chain (the pivot factorisation) then never queues behind the 16-cycle DMMAs
    // of a tile warp, which slows it four-fold (measured).  Slots left over exit at once.
    const int hwarp = (int)threadIdx.x >> 5, gid = hwarp / hw_per_group, hl = hwarp % hw_per_group;
    const int ntw = (gthreads >> 5) - (HELPER ? 1 : 0);
    int role = hl;
    if (HELPER) {
        role = ((hl & 3) == 3) ? ((hl == 3) ? ntw : -1) : (hl >> 2) * 3 + (hl & 3);
        if (role > ntw || ((hl & 3) != 3 && role >= ntw)) role = -1;
    }
    const int e = blockIdx.x * nmat + gid;
    if (role < 0 || e >= b.M) return;   // spare slots; a whole group without work: its named barriers are never used
    const int tid = role * 32 + ((int)threadIdx.x & 31);
    const bool next_group_exists = (gid + 1 < nmat) && (e + 1 < b.M);
    double* smem = smem_all + (size_t)gid * smem_doubles_per_group;
    const int lane = tid & 31, warp = tid >> 5, nwarps = ntw;   // tile warps; with HELPER, warp `nwarps` is the helper
    const bool helper = HELPER && (warp == nwarps);
    const int ntiles = T * (T + 1) / 2;
    const int lr = lane >> 2, lc = (lane & 3) * 2;   // this lane's row and first column inside a tile

    const unsigned pb = 512u * Tpad;                                // bytes of one panel buffer (Tpad: T rounded up to 2 nwarps)
    const unsigned gb = (unsigned)__cvta_generic_to_shared(smem);   // shared-window address of this group
    const unsigned lane16 = 16u * lane;
    const unsigned ltr = 16u * (lc * 4 + (lr >> 1)) + 8u * (lr & 1);   // this lane's first element in the TRANSPOSED tile (second: + 64)
    const unsigned opark = O_TM + (unsigned)nwarps * (SLOTS - RS) * 512;   // (HELPER = false) [RS] tiles of the factoring warp
    const unsigned odesc = opark + (HELPER ? 0u : (unsigned)RS * 512u), otab = odesc + (unsigned)nwarps * 256;
    const unsigned ozb = otab + (unsigned)nwarps * 256;
    const unsigned tm = gb + O_TM + (unsigned)warp * (SLOTS - RS) * 512 + lane16;   // this lane's pair of the warp's first memory tile
    const unsigned dw = gb + odesc + (unsigned)warp * 256, tw = gb + otab + (unsigned)warp * 256;
    double* Wb = smem + O_W / 8;        // [64] W = L^-1 row major, [64] W' row major
    double* misc = smem + O_MISC / 8;   // [0] quadratic form, [2] info (as int)
    double* red = smem + O_RED / 8;     // [16] warp totals
    double* rv = smem + O_RV / 8;       // [8 Tpad] residual, swept along with the matrix: ends as a = K~^-1 r
    double* piv = smem + O_PIV / 8;     // [Np] Schur pivots
    double* Zb = smem + ozb / 8;        // [Tpad][64] Z, then -Z, then [2][Tpad][64] panel C / X = C D^-1
    double2* stage = reinterpret_cast<double2*>(Zb);   // [nwarps][8][32] fragment staging (assembly, gradient): aliases the panels
    const size_t panel_doubles = (size_t)4 * Tpad * 64, stage_doubles = (size_t)nwarps * 512;
    double* tsh = Zb + (panel_doubles > stage_doubles ? panel_doubles : stage_doubles);   // [Np] shifted times
    double* av = tsh + Np;              // [Np] alpha per point (0 on padding)
    double* sbv = av + Np;              // [Np] Sigma_b per point
    double* dadd = sbv + Np;            // [Np] sigma^2 (1 on padding: identity pivots)
    // [T][T][8] + [T][8] gradient partials: behind the staging buffer inside the (by then idle) panel region when they fit
    const size_t part_doubles = b.want_grad ? (size_t)T * T * 8 + (size_t)T * 8 : 0;
    const bool part_aliased = stage_doubles + part_doubles <= panel_doubles;
    double* part = part_aliased ? Zb + stage_doubles : dadd + Np;
    double* partd = part + (b.want_grad ? T * T * 8 : 0);
    int* bandv = reinterpret_cast<int*>(dadd + Np + (part_aliased ? 0 : part_doubles));   // [Np]
#ifdef GPCC_FRAG_PROF
    long long* tlbuf = reinterpret_cast<long long*>(smem_all + (size_t)nmat * smem_doubles_per_group);
#endif
    double2* st = stage + warp * 256;

    const double rho = b.rho[e];
    const KernParams kp = make_kern_params(KID, rho);

    for (int i = tid; i < 8 * Tpad; i += gthreads) {
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
        rv[i] = r;
        if (i < Np) { tsh[i] = ts; av[i] = al; sbv[i] = sb; dadd[i] = dd; bandv[i] = bi; }
    }
    for (int i = tid; i < 128; i += gthreads) Wb[i] = 0.0;
    if (tid == 0) misc[0] = 0.0;
    // Slot s of this warp is tile q = ((s / CHUNK) * nwarps + warp) * CHUNK + s % CHUNK (chunks of consecutive tiles dealt
    // round robin); lane s keeps that tile's coordinates for the whole warp and writes its entry of the warp's
    // update-operand table {512 ti, 512 tj}.
    int mytile = -1;
    {
        const int q = ((lane / CHUNK) * nwarps + warp) * CHUNK + lane % CHUNK;
        if (!helper && q < ntiles) { int a_, b_; tile_of(q, a_, b_); mytile = (a_ << 8) | b_; }
        if (!helper) sts64u(tw + lane * 8, mytile < 0 ? 0u : (unsigned)(mytile >> 8) * 512u, mytile < 0 ? 0u : (unsigned)(mytile & 255) * 512u);
    }
    // bit s: the A operand (tile row) of slot s differs from that of slot s-1
    const int prevtile = __shfl_up_sync(0xffffffffu, mytile, 1);
    const unsigned amask = __ballot_sync(0xffffffffu, (mytile >> 8) != (prevtile >> 8)) | 1u;
    group_sync(gid, gthreads);

    // ---- assembly, eight tiles at a time through the staging buffer (keeps the exp code out of the unrolled part) ---
    double acc[RS][2];
    if (!helper)   // (the helper warp has no tiles: its accumulators stay undefined and are never read)
#pragma unroll
    for (int c = 0; c < SLOTS / 8; ++c) {
        for (int u = 0; u < 8; ++u) {
            const int sl = c * 8 + u;
            const int q = ((sl / CHUNK) * nwarps + warp) * CHUNK + sl % CHUNK;
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
    if (!helper) {
        write_descriptors(dw + lane * 8, mytile, -1, 0u, gb + ozb + 2 * pb, gb);   // panel 0, pivot tile 0 -> Db, pivot tile 1 -> Dn
        __syncwarp();
        scan_tiles(acc, tm, dw, lane16);
        if (!HELPER && warp == 0) {   // tile (0,0) is slot 0 of warp 0
            __syncwarp();
            if (lane == 0) pivot_tile(gb, 0);
        }
    }
    if (HELPER) group_sync(gid, gthreads);
    int pw = (nwarps > 1) ? 1 : 0;   // (HELPER = false) warp that owns block k+1 of the panel in step k: (k+1) % nwarps
    if (helper) {
        // Look-ahead: while the tile warps run the trailing update of step k, the helper warp finishes the next pivot
        // tile (D - Z Z' with Z = block k+1 of this step's panel, two DMMAs) and ONE of its lanes factors it: Cholesky of
        // the 8x8 block, W = L^-1, zr = W r_{k+1}.  That chain needs the registers of a whole thread and ~2 k cycles of
        // dependent FP64 latency; on a warp that also owns tiles it would be the critical path of every step.
        if (lane == 0) pivot_tile(gb, 0);
        for (int k = 0; k < T; ++k) {
            PROF_TL(0);
            group_sync(gid, gthreads);   // B1
            PROF_TL(1); PROF_TL(2);
            group_sync(gid, gthreads);   // B2
            PROF_TL(3);
            if (k == 0 && next_group_exists) asm volatile("bar.arrive %0, %1;" ::"r"(8 + gid + 1), "r"(2 * gthreads) : "memory");
            if (k + 1 < T) {
                const double2 z = lds128(gb + ozb + (k + 1) * 512 + lane16);
                double2 d = lds128(gb + O_DN + ((k + 1) & 1) * 512 + lane16);
                dmma884(d.x, d.y, negd(z.x), z.x);
                dmma884(d.x, d.y, negd(z.y), z.y);
                sts128(gb + O_DB + lane16, d.x, d.y);
                __syncwarp();
                if (lane == 0) pivot_tile(gb, k + 1);
            }
            PROF_TL(4); PROF_TL(5); PROF_TL(6);
        }
    } else
    for (int k = 0; k < T; ++k) {
        const unsigned cw = gb + ozb + (2 + (k & 1)) * pb, cn = gb + ozb + (3 - (k & 1)) * pb;
        PROF_TL(0);
        group_sync(gid, gthreads);
        PROF_TL(1);
        // P3: Z = C W', X = Z W, r -= Z zr on DMMA, two independent 8-row blocks of the panel in flight per warp.
        // Block t of the panel holds C rows 8t..8t+7 in the fragment layout of the tile it came from (transposed for
        // t < k: those come from row tiles), and receives X in the layout of the tile that will read it.
        {
            const double2 wz = lds128(gb + O_W + lane16);     // W[n=lr][2m, 2m+1]
            const double2 wx = lds128(gb + O_WT + lane16);    // W[2m, 2m+1][n=lr]
            const double2 zq = lds128(gb + O_ZR + (lane16 & 48));
            for (int t0 = warp; t0 < Tpad; t0 += 2 * nwarps) {
                // no branches in here: a lone warp pays ~30 cycles for every taken one
                double z0[2], z1[2], x0[2], x1[2], rold[2], dot[2];
                double2 c[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = t0 + u * nwarps;
                    const unsigned blk = cw + t * 512;
                    const unsigned off0 = (t < k) ? ltr : lane16, off1 = off0 + ((t < k) ? 64u : 8u);
                    c[u].x = lds64(blk + off0);
                    c[u].y = lds64(blk + off1);
                    rold[u] = lds64(gb + O_RV + t * 64 + lr * 8);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {   // pivot block: identity rows, X = D^-1
                    const bool pv = (t0 + u * nwarps == k);
                    c[u].x = pv ? ((lr == lc) ? 1.0 : 0.0) : c[u].x;
                    c[u].y = pv ? ((lr == lc + 1) ? 1.0 : 0.0) : c[u].y;
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) { z0[u] = 0.0; z1[u] = 0.0; dmma884(z0[u], z1[u], c[u].x, wz.x); }
#pragma unroll
                for (int u = 0; u < 2; ++u) dmma884(z0[u], z1[u], c[u].y, wz.y);
#pragma unroll
                for (int u = 0; u < 2; ++u) { x0[u] = 0.0; x1[u] = 0.0; dmma884(x0[u], x1[u], z0[u], wx.x); }
#pragma unroll
                for (int u = 0; u < 2; ++u) dmma884(x0[u], x1[u], z1[u], wx.y);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const unsigned za = gb + ozb + (t0 + u * nwarps) * 512 + lane16;
                    sts128(za, z0[u], z1[u]);                       // Zb
                    sts128(za + pb, negd(z0[u]), negd(z1[u]));      // Zn
                    dot[u] = fma(z1[u], zq.y, z0[u] * zq.x);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], 1);
#pragma unroll
                for (int u = 0; u < 2; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], 2);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = t0 + u * nwarps;
                    const unsigned blk = cw + t * 512;
                    const unsigned off0 = (t < k) ? ltr : lane16, off1 = off0 + ((t < k) ? 64u : 8u);
                    const int sg = (t == k) ? (int)0x80000000 : 0;   // pivot tile: -D^-1
                    sts64(blk + off0, __hiloint2double(__double2hiint(x0[u]) ^ sg, __double2loint(x0[u])));
                    sts64(blk + off1, __hiloint2double(__double2hiint(x1[u]) ^ sg, __double2loint(x1[u])));
                    const double rnew = (t == k) ? dot[u] : rold[u] - dot[u];   // r_k <- W' zr = D^-1 r_k
                    if ((lane & 3) == 0) sts64(gb + O_RV + t * 64 + lr * 8, rnew);
                }
            }
        }
        PROF_TL(2);
        group_sync(gid, gthreads);
        PROF_TL(3);
        if (k == 0 && next_group_exists) asm volatile("bar.arrive %0, %1;" ::"r"(8 + gid + 1), "r"(2 * gthreads) : "memory");
        write_descriptors(dw + lane * 8, mytile, k, cw, cn, gb);   // read back after the update (same warp: __syncwarp)
        if (!HELPER) {
            // Look-ahead without a helper warp: the tile warp that owns block k+1 of the panel finishes the next pivot tile
            // (D - Z Z') and ONE of its lanes factors it while the other warps run their update; the factorisation needs
            // the registers of a whole thread, so the warp's tiles are parked in shared memory meanwhile.
            if (k + 1 < T && warp == pw) {
                const double2 z = lds128(gb + ozb + (k + 1) * 512 + lane16);
                double2 d = lds128(gb + O_DN + ((k + 1) & 1) * 512 + lane16);
                dmma884(d.x, d.y, negd(z.x), z.x);
                dmma884(d.x, d.y, negd(z.y), z.y);
                sts128(gb + O_DB + lane16, d.x, d.y);
#pragma unroll
                for (int s = 0; s < RS; ++s) sts128(gb + opark + s * 512 + lane16, acc[s][0], acc[s][1]);
                __syncwarp();
                if (lane == 0) pivot_tile(gb, k + 1);
#pragma unroll
                for (int s = 0; s < RS; ++s) { const double2 v = lds128(gb + opark + s * 512 + lane16); acc[s][0] = v.x; acc[s][1] = v.y; }
            }
            pw = (pw + 1 == nwarps) ? 0 : pw + 1;
        }
        PROF_TL(4);
        // bulk: A_ij -= Z_i Z_j' on every tile of the warp (tiles of row / column k are overwritten right after).
        // Operand addresses come from the warp's table; everything is fetched one tile ahead (a warp issues in order).
        {
            const unsigned zbb = gb + ozb + lane16, znb = zbb + pb;
            uint2 o = lds64u(tw);
            double2 a_0 = lds128(znb + o.x), b_0 = lds128(zbb + o.y);
            uint2 o1 = lds64u(tw + 8);
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                double2 a_1 = a_0, b_1 = b_0;
                uint2 o2 = o1;
                if (s + 1 < SLOTS) {
                    if ((amask >> (s + 1)) & 1u) a_1 = lds128(znb + o1.x);
                    b_1 = lds128(zbb + o1.y);
                    if (s + 2 < SLOTS) o2 = lds64u(tw + (s + 2) * 8);
                }
                if (s < RS) {
                    dmma884(acc[s < RS ? s : 0][0], acc[s < RS ? s : 0][1], a_0.x, b_0.x);
                    dmma884(acc[s < RS ? s : 0][0], acc[s < RS ? s : 0][1], a_0.y, b_0.y);
                } else {
                    double2 d = lds128(tm + (s - RS) * 512);
                    dmma884(d.x, d.y, a_0.x, b_0.x);
                    dmma884(d.x, d.y, a_0.y, b_0.y);
                    sts128(tm + (s - RS) * 512, d.x, d.y);
                }
                a_0 = a_1; b_0 = b_1; o1 = o2;
            }
        }
        PROF_TL(5);
        // scan: final values of row / column k, next panel, raw pivot tile k+2
        __syncwarp();
        scan_tiles(acc, tm, dw, lane16);
        PROF_TL(6);
    }
    group_sync(gid, gthreads);
#ifdef GPCC_FRAG_PROF
    if (blockIdx.x == 0) for (int i = tid; i < T * 8 * 16; i += gthreads) if ((i % 16) / (nwarps + 1) == gid) g_frag_tl[i] = tlbuf[i];
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
    if (!helper)
#pragma unroll
    for (int c = 0; c < SLOTS / 8; ++c) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { double g0, g1; TILE_GET(c * 8 + u, g0, g1); st[u * 32 + lane] = make_double2(g0, g1); }
        __syncwarp();
        for (int u = 0; u < 8; ++u) {
            const int sl = c * 8 + u;
            const int q = ((sl / CHUNK) * nwarps + warp) * CHUNK + sl % CHUNK;
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
    for (int pb = warp; pb < L; pb += (gthreads >> 5)) {
        double s = 0.0;
        for (int i = p.band_start[pb] + lane; i < p.band_start[pb + 1]; i += 32) s += srow[i];
        s = warp_sum(s);
        if (lane == 0) b.grad[(size_t)e * (L + 1) + pb] = s / b.alpha[(size_t)e * L + pb];
    }
    if (tid == 0) b.grad[(size_t)e * (L + 1) + L] = es;   // 0.5 * sum_full = sum over the strict lower triangle
}

int padded_T(int T, int nwarps) { return (T + 2 * nwarps - 1) / (2 * nwarps) * (2 * nwarps); }

template <bool HELPER>
size_t group_smem_doubles(int T, int nwarps, int want_grad) {
    constexpr int RS = FragCfg<HELPER>::RS;
    const int Np = 8 * T, Tpad = padded_T(T, nwarps);
    const size_t panel = (size_t)4 * Tpad * 64, stage = (size_t)nwarps * 512;
    const size_t part = want_grad ? (size_t)T * T * 8 + (size_t)T * 8 : 0;
    size_t doubles = (O_TM + (size_t)nwarps * ((SLOTS - RS) * 512 + 512) + (HELPER ? 0 : (size_t)RS * 512)) / 8 + (panel > stage ? panel : stage) +
                     (size_t)Np * 4 + (stage + part <= panel ? 0 : part) + (size_t)(Np + 1) / 2 + 2;
    return (doubles + 1) & ~(size_t)1;   // keep every group 16-byte aligned
}

template <int KID, bool HELPER>
cudaError_t launch_cfg(const DevProblem& p, const EvalBatch& b, int T, cudaStream_t s) {
    const int ntiles = T * (T + 1) / 2;
    const int nwarps = (ntiles + SLOTS - 1) / SLOTS;
    const int gthreads = (nwarps + (HELPER ? 1 : 0)) * 32;
    static const int nmat_cap = getenv("GPCC_FRAG_NMAT") ? atoi(getenv("GPCC_FRAG_NMAT")) : (HELPER ? 4 : 1);
    const size_t gd = group_smem_doubles<HELPER>(T, nwarps, b.want_grad);
    const int hwpg = HELPER ? 4 * ((nwarps + 2) / 3) : nwarps;   // hardware warps per matrix (HELPER: three tile warps per quad + the helper / spare slot)
    int nmat = ((HELPER ? MAX_THREADS : 384) / 32) / hwpg;
    if (nmat > nmat_cap) nmat = nmat_cap;
    while (nmat > 1 && gd * 8 * nmat > 220 * 1024) --nmat;
    if (nmat < 1) nmat = 1;
    auto kfn = small_frag_kernel<KID, HELPER>;
    size_t extra = 0;
#ifdef GPCC_FRAG_PROF
    extra = (size_t)T * 8 * 16 * 8;
#endif
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_done = true; }
    const int blocks = (b.M + nmat - 1) / nmat;
    kfn<<<blocks, hwpg * 32 * nmat, gd * 8 * nmat + extra, s>>>(p, b, T, padded_T(T, nwarps), gthreads, hwpg, nmat, (int)gd);
    cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kfn);
        fprintf(stderr, "[small_frag] launch failed: %s; threads %d, dynamic smem %zu B, kernel: %d regs, %zu B static smem, max %d threads/block, max dynamic smem %d\n",
                cudaGetErrorString(rc), hwpg * 32 * nmat, gd * 8 * nmat + extra, fa.numRegs, fa.sharedSizeBytes, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes);
    }
    return rc;
}

template <int KID>
cudaError_t launch_kid(const DevProblem& p, const EvalBatch& b, int T, cudaStream_t s) {
    static const int helper = getenv("GPCC_FRAG_HELPER") ? atoi(getenv("GPCC_FRAG_HELPER")) : 0;
    return helper ? launch_cfg<KID, true>(p, b, T, s) : launch_cfg<KID, false>(p, b, T, s);
}

}  // namespace

bool small_frag_supports(int N) { return (N + 7) / 8 <= MAX_T; }

cudaError_t small_frag_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t s) {
    const int T = (p.N + 7) / 8;
#ifdef GPCC_FRAG_PROF
    cudaError_t rc = launch_kid<K_M32>(p, b, T, s);
    cudaStreamSynchronize(s);
    static long long tl[32 * 8 * 16];
    cudaMemcpyFromSymbol(tl, g_frag_tl, sizeof(tl));
    const int nw = ((T * (T + 1) / 2 + 31) / 32);
    for (int g = 0; g < 14 / (nw + 1) && g < 4; ++g) {
        // slots: 0 step start, 1 after B1, 2 P3 done, 3 after B2, 4 look-ahead pivot done, 5 bulk done, 6 scan done
        double acc[6] = {0}; long long first = 0, last = 0;
        for (int k = 1; k < T; ++k) {
            long long mn0 = 1LL << 62, mx[8] = {0};
            for (int w = 0; w <= nw; ++w) {
                const long long* q = tl + (k * 8) * 16 + w + (nw + 1) * g;
                mn0 = q[0] < mn0 ? q[0] : mn0;
                for (int sl = 0; sl < 7; ++sl) mx[sl] = q[16 * sl] > mx[sl] ? q[16 * sl] : mx[sl];
            }
            if (k == 1) first = mn0;
            last = mx[6];
            acc[0] += mx[1] - mn0; acc[1] += mx[3] - mx[1]; acc[2] += mx[4] - mx[3]; acc[3] += mx[5] - mx[4]; acc[4] += mx[6] - mx[5];
        }
        if (last == 0) continue;
        if (g == 0) for (int w = 0; w <= nw; ++w) {
            const long long* q = tl + (5 * 8) * 16 + w;
            fprintf(stderr, "[frag tl] T=%d k=5 warp %d: start %lld | after B1 %lld | P3 done %lld | after B2 %lld | pivot done %lld | bulk done %lld | scan done %lld\n", T, w,
                    q[0] - tl[5 * 8 * 16], q[16] - tl[5 * 8 * 16], q[32] - tl[5 * 8 * 16], q[48] - tl[5 * 8 * 16], q[64] - tl[5 * 8 * 16], q[80] - tl[5 * 8 * 16], q[96] - tl[5 * 8 * 16]);
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
