// Fused small-N evaluator: one CTA per (delay candidate, hyper-parameter) pair.
//
// Replaces, for N+1 <= 8*SMALL_MAX_T, the reference's objective
//   K = delayedCovariance(kernel, alpha, tau, rho, tarray) + Sobs + B ; logpdf(MvNormal(bbar, K), Y)
// (/root/reference/src/gpccfixdelay_marginaliseb.jl:133-141, src/delayedCovariance.jl:1-38) and adds
// the analytic gradient 0.5 tr((a a' - K^-1) dK/dtheta) that north_star asks for.
//
// Design (B200: 64 FP64 FMA/clk/SM, 64K registers/SM, smem 128 B/clk):
//   * the lower triangle of the (N+1)x(N+1) bordered matrix [K~ r; r' 0] lives ENTIRELY IN REGISTERS,
//     one 8x8 tile per thread (N=150 -> 190 threads x 64 doubles); nothing N^2-sized touches smem/HBM;
//   * assembly is fused: every thread evaluates its 64 kernel entries from the shifted times in smem;
//   * one symmetric Gauss-Jordan "sweep" per index k (k = 0..N-1): A_ij -= A_ik A_kj / A_kk for
//     i,j != k, A_ik <- A_ik/A_kk, A_kk <- -1/A_kk.  Per step a thread does 64 independent DFMAs on its
//     tile from 16 values broadcast through shared memory (8 LDS.128) and ONE __syncthreads: the
//     owners publish column k+1 (double buffered) as soon as their slice of it is updated.
//     The pivots are exactly the Cholesky pivots L_kk^2, so logdet = sum log(pivot) and the
//     leading-minor `info` follow LAPACK dpotrf; the border row gives a = K~^-1 r and the corner
//     -r'K~^-1 r by forward elimination; after N steps the tile registers hold -K~^-1.
//     Work: N^3/2 DFMA = N^3 flop, the same as potrf + potri, but perfectly balanced, with no
//     triangular solves and no second pass;
//   * the gradient contracts W = a a' - K~^-1 against K and dK/drho recomputed on the fly (never
//     stored), reduced deterministically (no atomics).
#include "gpcc_internal.h"
#include "kernfun.cuh"
#include <climits>
#include <cmath>
#include <cstdlib>

namespace gpcc {

namespace {

constexpr int TS = SMALL_TILE;  // 8

// Two matrices may share one CTA (NMAT = 2): each "group" of threads owns one matrix, its own slice of shared memory and
// its own named barrier.  With 6 warps per matrix, warps 0-5 and 6-11 of one CTA land on the four schedulers 3 + 3 + 3 + 3;
// two separate CTAs of 6 warps land 4 + 4 + 2 + 2, i.e. two schedulers carry twice the DFMA load of the others.
#define GROUP_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(nthreads) : "memory")
constexpr double LOG2PI = 1.8378770664093454835606594728112;

// Chunked layout of the per-point vectors in shared memory: [pair-of-rows part (4)][tile (<=32)][2], with a
// compile-time part stride so that every address in the sweep loop is "per-thread base + immediate".
constexpr int CS = 64;          // doubles per part (>= 2*SMALL_MAX_T)
constexpr int VLEN = 4 * CS;    // doubles per vector
__device__ __forceinline__ int cidx(int i, int /*T*/) {
    return ((i & 7) >> 1) * CS + ((i >> 3) << 1) + (i & 1);
}
__device__ __forceinline__ void load8(const double* buf, int tile, int T, double (&out)[8]) {
#pragma unroll
    for (int part = 0; part < 4; ++part) {
        const double2 v = *reinterpret_cast<const double2*>(buf + part * CS + 2 * tile);
        out[2 * part] = v.x;
        out[2 * part + 1] = v.y;
    }
}
__device__ __forceinline__ void store8(double* buf, int tile, int T, const double (&in)[8]) {
#pragma unroll
    for (int part = 0; part < 4; ++part)
        *reinterpret_cast<double2*>(buf + part * CS + 2 * tile) = make_double2(in[2 * part], in[2 * part + 1]);
}

// Publish column `kn` (tile tkn, in-tile index KKN) of the symmetric matrix for the next sweep step.
// Scalar STS.64 straight from the tile registers (no staging moves); only warps that contain an owner enter.
template <int KKN, bool FWD = false>
__device__ __forceinline__ void publish(const double (&A)[8][8], int ti, int tj, int tkn, int kn, int T,
                                        double* nb, double* pslot, double* piv, bool active) {
    // forward-only mode: the row part (tiles left of the pivot tile) is never read again
    const bool pc = active && (tj == tkn), prw = active && (ti == tkn) && (!FWD || tj == tkn);
    if (!__any_sync(0xffffffffu, pc || prw)) return;
    if (pc) {
        double* dst = nb + 2 * ti;
        if (prw) {  // diagonal tile: below the diagonal from the column, above it from the row
#pragma unroll
            for (int r = 0; r < 8; ++r) dst[(r >> 1) * CS + (r & 1)] = (r >= KKN) ? A[r][KKN] : A[KKN][r];
            const double d = A[KKN][KKN];
            piv[kn] = d;
            *pslot = 1.0 / d;
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) dst[(r >> 1) * CS + (r & 1)] = A[r][KKN];
        }
    } else if (prw) {
        double* dst = nb + 2 * tj;
#pragma unroll
        for (int c = 0; c < 8; ++c) dst[(c >> 1) * CS + (c & 1)] = A[KKN][c];
    }
}

// FWD (forward-only mode, no gradient): plain Gaussian elimination.  Only the tiles with tj >= tk still change something that
// is read later (the pivots and the border row); the others skip the step, so the work is N^3/3 flop instead of N^3.  The
// elements that feed the pivots and the corner see exactly the operations of the full sweep: logL is bitwise the same.
template <int KK, bool FWD>
__device__ __forceinline__ void sweep_step(double (&A)[8][8], int ti, int tj, int tk, int k, int N, int T, int Np,
                                           double* cbuf, double* pbuf, double* piv, bool active, int gid, int nthreads) {
    if (FWD && tj < tk) { GROUP_SYNC(); return; }
    const double* cb = cbuf + (k & 1) * VLEN;
    double v[8];
    load8(cb, tj, T, v);
    const double pr = pbuf[k & 1];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] *= pr;
    // Column k of its owner tiles must become c * (1/d).  Those registers hold exactly the broadcast values c, so
    // running the generic update with the multiplier (1 - 1/d) in that column writes c - c (1 - 1/d) = c/d without
    // any extra instruction (relative error eps*d in entries of the inverse only; pivots, log-det and the
    // quadratic form never read the swept region).
    // Row k of its owner tiles must become c * (1/d) as well: there the registers hold c and the row multiplier is the
    // pivot d itself, so using (d - 1) instead gives c - (d - 1) c/d = c/d.  Only the diagonal element needs a fix.
    const bool own_col = (tj == tk), own_row = (ti == tk);
    if (own_col) v[KK] = 1.0 - pr;
#pragma unroll
    for (int part = 0; part < 4; ++part) {   // rows two at a time: keeps only a pair of broadcast values live
        double2 x = *reinterpret_cast<const double2*>(cb + part * CS + 2 * ti);
        if (part == (KK >> 1) && own_row) {
            if (KK & 1) x.y -= 1.0; else x.x -= 1.0;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            A[2 * part][c] = fma(-x.x, v[c], A[2 * part][c]);
            A[2 * part + 1][c] = fma(-x.y, v[c], A[2 * part + 1][c]);
        }
    }
    if (own_col && own_row) A[KK][KK] = -pr;
    const int kn = k + 1;
    if (kn < N) {
        double* nb = cbuf + (kn & 1) * VLEN;
        if (KK < 7) publish<(KK + 1) & 7, FWD>(A, ti, tj, tk, kn, T, nb, pbuf + (kn & 1), piv, active);
        else        publish<0, FWD>(A, ti, tj, tk + 1, kn, T, nb, pbuf + (kn & 1), piv, active);
    }
    GROUP_SYNC();
}

// ---- two pivots per barrier (PAIR = true, even N) -----------------------------------------------------------------------
// Columns k and k+1 of the symmetric matrix are published together, as they stand before step k.  Every thread rebuilds
// the entries of column k+1 it needs AFTER step k itself (c'_i = c_{k+1,i} - c_{k,i} m with m = c_{k,k+1} / d_k: the very
// FMA the owner of that entry performs in step k, so the values are identical), forms d'_{k+1} = d_{k+1} - c_{k,k+1} m and
// its reciprocal, and applies the two rank-1 steps back to back, two tile rows at a time.  The arithmetic is that of the
// rank-1 kernel (no 2x2 pivot block is inverted); what halves is everything paid once per barrier: the barrier itself, the
// shared-memory round trip behind it, the divergent owner paths of `publish` and their branches, and the reciprocal chain.
template <int KKN>
__device__ __forceinline__ void publish2(const double (&A)[8][8], int ti, int tj, int tkn, int kn, double* nb, double* pslot,
                                         double* piv, bool active) {
    const bool pc = active && (tj == tkn), prw = active && (ti == tkn);
    if (!__any_sync(0xffffffffu, pc || prw)) return;
    if (pc) {
        double* da = nb + 2 * ti;
        double* db = da + VLEN;
        if (prw) {  // diagonal tile: below the diagonal from the column, above it from the row
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                da[(r >> 1) * CS + (r & 1)] = (r >= KKN) ? A[r][KKN] : A[KKN][r];
                db[(r >> 1) * CS + (r & 1)] = (r >= KKN + 1) ? A[r][KKN + 1] : A[KKN + 1][r];
            }
            const double d = A[KKN][KKN];
            piv[kn] = d;
            *pslot = 1.0 / d;
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                da[(r >> 1) * CS + (r & 1)] = A[r][KKN];
                db[(r >> 1) * CS + (r & 1)] = A[r][KKN + 1];
            }
        }
    } else if (prw) {
        double* da = nb + 2 * tj;
        double* db = da + VLEN;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            da[(c >> 1) * CS + (c & 1)] = A[KKN][c];
            db[(c >> 1) * CS + (c & 1)] = A[KKN + 1][c];
        }
    }
}

template <int KK>   // KK even: pivots k = 8 tk + KK and k + 1
__device__ __forceinline__ void pair_step(double (&A)[8][8], int ti, int tj, int tk, int k, int N, double* cbuf, double* pbuf,
                                          double* piv, bool active, int tid, int gid, int nthreads) {
    const int par = (k >> 1) & 1;
    const double* ca = cbuf + par * 2 * VLEN;
    const double* cq = ca + VLEN;
    double vA[8], vB[8];
    load8(ca, tj, 0, vA);
    load8(cq, tj, 0, vB);
    const double pr = pbuf[par];
    const int ik1 = ((KK + 1) >> 1) * CS + tk * 2 + 1;     // chunk index of row k + 1
    const double ckk1 = ca[ik1], dB = cq[ik1];
    const double m = ckk1 * pr;
    const double dBp = fma(-ckk1, m, dB);                   // pivot k+1 after step k
    const double prB = 1.0 / dBp;
    if (tid == 0) piv[k + 1] = dBp;
    const bool own_col = (tj == tk), own_row = (ti == tk);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double xa = vA[c];
        if (c == KK) xa = own_col ? xa - 1.0 : xa;         // the row-k value of step k for index (tj, c)
        vB[c] = fma(-xa, m, vB[c]) * prB;                   // column k+1 after step k, scaled by 1/d'
        vA[c] *= pr;
    }
    vA[KK] = own_col ? 1.0 - pr : vA[KK];
    vB[KK + 1] = own_col ? 1.0 - prB : vB[KK + 1];
#pragma unroll
    for (int part = 0; part < 4; ++part) {
        double2 xa = *reinterpret_cast<const double2*>(ca + part * CS + 2 * ti);
        const double2 xb = *reinterpret_cast<const double2*>(cq + part * CS + 2 * ti);
        if (part == (KK >> 1)) xa.x = own_row ? xa.x - 1.0 : xa.x;
        double2 xp;
        xp.x = fma(-xa.x, m, xb.x);
        xp.y = fma(-xa.y, m, xb.y);
        if (part == (KK >> 1)) xp.y = own_row ? xp.y - 1.0 : xp.y;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            A[2 * part][c] = fma(-xa.x, vA[c], A[2 * part][c]);
            A[2 * part + 1][c] = fma(-xa.y, vA[c], A[2 * part + 1][c]);
        }
        if (part == (KK >> 1)) A[KK][KK] = (own_col && own_row) ? -pr : A[KK][KK];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            A[2 * part][c] = fma(-xp.x, vB[c], A[2 * part][c]);
            A[2 * part + 1][c] = fma(-xp.y, vB[c], A[2 * part + 1][c]);
        }
        if (part == (KK >> 1)) A[KK + 1][KK + 1] = (own_col && own_row) ? -prB : A[KK + 1][KK + 1];
    }
    const int kn = k + 2;
    if (kn < N) {
        double* nb = cbuf + (par ^ 1) * 2 * VLEN;
        if (KK < 6) publish2<(KK + 2) & 7>(A, ti, tj, tk, kn, nb, pbuf + (par ^ 1), piv, active);
        else        publish2<0>(A, ti, tj, tk + 1, kn, nb, pbuf + (par ^ 1), piv, active);
    }
    GROUP_SYNC();
}

// ---- branch-free form of the step (BF = true) -------------------------------------------------------------------------
// Measured on B200 (profiles/README.md, round-1 session 2): a warp issues in order and a lone warp pays ~25-30 cycles for
// every taken branch and ~30 for every shared-memory round trip, so with two or three warps per scheduler the eight
// branches of the step above (loop exits, the vote around `publish`, its divergent owner paths) are a fifth of the
// step.  Here the step is straight-line code: predicated stores for the owners (inline PTX, so that they cannot turn
// back into branches), a Newton reciprocal under the same predicate instead of the IEEE division, 32-bit shared-window
// addresses (one register per pointer instead of two: fewer spills at 168 registers), and no exit tests inside full tiles.
__device__ __forceinline__ double2 lds_v2(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_pred(unsigned a, double v, int flag) {
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; @p st.shared.f64 [%0], %1; }" ::"r"(a), "d"(v), "r"(flag) : "memory");
}
// pivot d of the next step -> piv[kn], 1/d -> pslot, only where flag is set (the thread that owns the diagonal element)
__device__ __forceinline__ void publish_pivot(unsigned piv_a, unsigned pslot_a, double d, int flag) {
    asm volatile(
        "{ .reg .pred p; .reg .f64 x, e, n;\n"
        "  setp.ne.s32 p, %3, 0;\n"
        "  @p rcp.approx.ftz.f64 x, %2;\n"
        "  @p neg.f64 n, %2;\n"
        "  @p fma.rn.f64 e, n, x, 0d3FF0000000000000;\n"
        "  @p fma.rn.f64 x, x, e, x;\n"
        "  @p fma.rn.f64 e, n, x, 0d3FF0000000000000;\n"
        "  @p fma.rn.f64 x, x, e, x;\n"
        "  @p fma.rn.f64 e, n, x, 0d3FF0000000000000;\n"
        "  @p fma.rn.f64 x, x, e, x;\n"
        "  @p st.shared.f64 [%0], %2;\n"
        "  @p st.shared.f64 [%1], x;\n"
        "}" ::"r"(piv_a), "r"(pslot_a), "d"(d), "r"(flag) : "memory");
}
// a_ti / a_tj: shared address of this thread's row / column pair slot in broadcast buffer 0 (cbuf + 2 ti, cbuf + 2 tj);
// buffers alternate with the step parity, VLEN doubles apart; parts (row pairs) are CS doubles apart.
template <int KKN>
__device__ __forceinline__ void publish_bf(const double (&A)[8][8], int pc, int prw, unsigned dst_ti, unsigned dst_tj,
                                           unsigned piv_a, unsigned pslot_a) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {   // column kn from its tile column; on the diagonal tile the part above the diagonal comes from the row
        const double val = (r < KKN) ? (prw ? A[KKN][r] : A[r][KKN]) : A[r][KKN];
        sts_pred(dst_ti + ((r >> 1) * CS + (r & 1)) * 8, val, pc);
    }
    const int ronly = prw & (pc ^ 1);
#pragma unroll
    for (int c = 0; c < 8; ++c) sts_pred(dst_tj + ((c >> 1) * CS + (c & 1)) * 8, A[KKN][c], ronly);
    publish_pivot(piv_a, pslot_a, A[KKN][KKN], pc & prw);
}
template <int KK, bool LAST>
__device__ __forceinline__ void sweep_step_bf(double (&A)[8][8], int ti, int tj, int tk, int k, unsigned a_ti, unsigned a_tj,
                                              unsigned pbuf_a, unsigned piv_a, int active, int gid, int nthreads) {
    constexpr unsigned BUF = (KK & 1) * VLEN * 8, NBUF = ((KK + 1) & 1) * VLEN * 8;
    double v[8];
#pragma unroll
    for (int part = 0; part < 4; ++part) {
        const double2 t = lds_v2(a_tj + BUF + part * CS * 8);
        v[2 * part] = t.x;
        v[2 * part + 1] = t.y;
    }
    const double pr = lds_f64(pbuf_a + (KK & 1) * 8);
    double2 x[4];
#pragma unroll
    for (int part = 0; part < 4; ++part) x[part] = lds_v2(a_ti + BUF + part * CS * 8);
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] *= pr;
    const bool own_col = (tj == tk), own_row = (ti == tk);
    v[KK] = own_col ? 1.0 - pr : v[KK];            // column k of its owners becomes c/d  (see sweep_step)
    {
        double& xk = (KK & 1) ? x[KK >> 1].y : x[KK >> 1].x;
        xk = own_row ? xk - 1.0 : xk;              // row k of its owners becomes c/d
    }
#pragma unroll
    for (int part = 0; part < 4; ++part) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            A[2 * part][c] = fma(-x[part].x, v[c], A[2 * part][c]);
            A[2 * part + 1][c] = fma(-x[part].y, v[c], A[2 * part + 1][c]);
        }
    }
    A[KK][KK] = (own_col && own_row) ? -pr : A[KK][KK];
    if (!LAST) {
        const int tkn = (KK < 7) ? tk : tk + 1;
        publish_bf<(KK + 1) & 7>(A, active & (tj == tkn), active & (ti == tkn), a_ti + NBUF, a_tj + NBUF, piv_a + (k + 1) * 8,
                                 pbuf_a + ((KK + 1) & 1) * 8);
    }
    GROUP_SYNC();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum: xor-tree inside each warp, then thread 0 adds the warp totals in order.
__device__ __forceinline__ double block_sum(double v, double* red, int tid, int nthreads, int gid) {
    v = warp_sum(v);
    GROUP_SYNC();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    GROUP_SYNC();
    double s = 0.0;
    const int nw = (nthreads + 31) >> 5;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

template <int KID, int MAXTHREADS, int MINBLOCKS, int NMAT, bool BF = false, bool PAIR = false, bool FWD = false>
__global__ void __launch_bounds__(MAXTHREADS, MINBLOCKS)
small_sweep_kernel(DevProblem p, EvalBatch b, int T, int group_doubles) {
    extern __shared__ __align__(16) double smem_all[];
    const int N = p.N, L = p.L;
    const int Np = T * TS;
    const int nthreads = blockDim.x / NMAT;
    const int gid = (NMAT == 1) ? 0 : (int)threadIdx.x / nthreads;
    const int tid = (int)threadIdx.x - gid * nthreads;
    const int e = blockIdx.x * NMAT + gid;
    if (e >= b.M) return;   // a whole group without work: its named barrier is never used
    double* smem = smem_all + (size_t)gid * group_doubles;
    const int ntiles = T * (T + 1) / 2;
    const bool active = tid < ntiles;
    const int q = active ? tid : ntiles - 1;
    int ti, tj;
    if (FWD) {
        // column-major tile order: the tiles that have left the elimination (tj < tk) are a prefix of the thread range, so
        // whole warps retire as the pivot moves on instead of every warp keeping a few live lanes
        int c = 0, rem = q;
        while (rem >= T - c) { rem -= T - c; ++c; }
        tj = c;
        ti = c + rem;
    } else {
        ti = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
        while (ti * (ti + 1) / 2 > q) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
        tj = q - ti * (ti + 1) / 2;
    }

    double* tsh = smem;             // shifted times t_i - tau_band(i)         (chunk layout)
    double* av = tsh + VLEN;        // alpha_band(i), 0 for padding            (chunk layout)
    double* sbv = av + VLEN;        // Sigma_b[band(i)]                         (chunk layout)
    double* dadd = sbv + VLEN;      // sigma_i^2                                (chunk layout)
    double* cbuf = dadd + VLEN;     // 2 x broadcast column                     (chunk layout)
    double* piv = cbuf + 4 * VLEN;  // pivots (cbuf: 2 buffers x 2 columns for the two-pivot step)   (natural)
    double* abuf = piv + VLEN;      // residual r, later a = K~^-1 r            (chunk layout)
    double* pbuf = abuf + VLEN;     // 2 pivot reciprocals (+2 pad)
    double* red = pbuf + 4;         // 64 reduction slots
    int* s_bad_p = reinterpret_cast<int*>(red + 48);
    double* part = red + 64;        // [T][T][8] gradient row-sum partials (gradient only)
    int* bandv = reinterpret_cast<int*>(part + (b.want_grad ? T * T * 8 : 0));  // [Np] natural

    const double rho = b.rho[e];
    const KernParams kp = make_kern_params(KID, rho);

    for (int i = tid; i < Np; i += nthreads) {
        const int ci = cidx(i, T);
        if (i < N) {
            const int bi = p.band[i];
            tsh[ci] = p.t[i] - b.delays[(size_t)e * L + bi];   // delayedCovariance.jl:27 (x - delays[l])
            av[ci] = b.alpha[(size_t)e * L + bi];
            sbv[ci] = b.mode_postb ? 0.0 : p.sigb[i];
            dadd[ci] = p.s2[i];
            abuf[ci] = b.mode_postb ? p.y[i] : p.resid[i];
            bandv[i] = bi;
        } else {
            tsh[ci] = 0.0; av[ci] = 0.0; sbv[ci] = 0.0; dadd[ci] = 0.0; abuf[ci] = 0.0;
            bandv[i] = -1 - i;
        }
    }
    GROUP_SYNC();

    // ---- assembly of the bordered matrix tile in registers --------------------------------------
    double A[8][8];
    {
        double tc[8], ac[8], rc[8];
        int bc[8];
        load8(tsh, tj, T, tc);
        load8(av, tj, T, ac);
        load8(abuf, tj, T, rc);
#pragma unroll
        for (int c = 0; c < 8; ++c) bc[c] = bandv[tj * 8 + c];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = ti * 8 + r;
            const int ci = cidx(i, T);
            const double tr = tsh[ci], ar = av[ci], sbr = sbv[ci], dr = dadd[ci];
            const int br = bandv[i];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int j = tj * 8 + c;
                const double kv = kern_value<KID>(tr - tc[c], kp);
                double val = (ar * ac[c]) * kv;          // scale[l]*scale[m]*kernel  (delayedCovariance.jl:27)
                if (i == j) val += dr;                   // + Sobs                   (gpccfixdelay_marginaliseb.jl:135)
                if (br == bc[c]) val += sbr;             // + B = Q Sigma_b Q'
                if (i == N) val = rc[c];                 // border row: r = Y - bbar (corner = 0)
                if (i > N && i == j) val = 1.0;          // padding
                A[r][c] = val;
            }
        }
    }
    GROUP_SYNC();   // everyone has read abuf/tsh before cbuf traffic starts (abuf is reused later)

    // ---- publish column 0, then N sweep steps ----------------------------------------------------
    if (PAIR) {   // N even (host dispatch)
        publish2<0>(A, ti, tj, 0, 0, cbuf, pbuf, piv, active);
        GROUP_SYNC();
        for (int tk = 0; tk < T; ++tk) {
            const int k0 = tk * 8;
            if (k0 >= N) break;
            pair_step<0>(A, ti, tj, tk, k0 + 0, N, cbuf, pbuf, piv, active, tid, gid, nthreads); if (k0 + 2 >= N) break;
            pair_step<2>(A, ti, tj, tk, k0 + 2, N, cbuf, pbuf, piv, active, tid, gid, nthreads); if (k0 + 4 >= N) break;
            pair_step<4>(A, ti, tj, tk, k0 + 4, N, cbuf, pbuf, piv, active, tid, gid, nthreads); if (k0 + 6 >= N) break;
            pair_step<6>(A, ti, tj, tk, k0 + 6, N, cbuf, pbuf, piv, active, tid, gid, nthreads);
        }
    } else {
    publish<0>(A, ti, tj, 0, 0, T, cbuf, pbuf, piv, active);
    GROUP_SYNC();
    if (BF) {
        const unsigned cb_a = (unsigned)__cvta_generic_to_shared(cbuf);
        const unsigned a_ti = cb_a + ti * 16, a_tj = cb_a + tj * 16;
        const unsigned pbuf_a = (unsigned)__cvta_generic_to_shared(pbuf), piv_a = (unsigned)__cvta_generic_to_shared(piv);
        const int act = active ? 1 : 0;
        const int full = N >> 3;   // tiles whose eight pivots are all < N: no exit tests inside (publishing column N is harmless)
        for (int tk = 0; tk < full; ++tk) {
            const int k0 = tk * 8;
            sweep_step_bf<0, false>(A, ti, tj, tk, k0 + 0, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            sweep_step_bf<1, false>(A, ti, tj, tk, k0 + 1, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            sweep_step_bf<2, false>(A, ti, tj, tk, k0 + 2, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            sweep_step_bf<3, false>(A, ti, tj, tk, k0 + 3, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            sweep_step_bf<4, false>(A, ti, tj, tk, k0 + 4, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            sweep_step_bf<5, false>(A, ti, tj, tk, k0 + 5, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            sweep_step_bf<6, false>(A, ti, tj, tk, k0 + 6, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            sweep_step_bf<7, false>(A, ti, tj, tk, k0 + 7, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
        }
        {   // the remaining N % 8 pivots (the border row N and the padding are never pivots)
            const int tk = full, k0 = full * 8, rem = N - k0;
            if (rem > 0) sweep_step_bf<0, false>(A, ti, tj, tk, k0 + 0, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            if (rem > 1) sweep_step_bf<1, false>(A, ti, tj, tk, k0 + 1, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            if (rem > 2) sweep_step_bf<2, false>(A, ti, tj, tk, k0 + 2, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            if (rem > 3) sweep_step_bf<3, false>(A, ti, tj, tk, k0 + 3, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            if (rem > 4) sweep_step_bf<4, false>(A, ti, tj, tk, k0 + 4, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            if (rem > 5) sweep_step_bf<5, false>(A, ti, tj, tk, k0 + 5, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
            if (rem > 6) sweep_step_bf<6, false>(A, ti, tj, tk, k0 + 6, a_ti, a_tj, pbuf_a, piv_a, act, gid, nthreads);
        }
    } else
    for (int tk = 0; tk < T; ++tk) {
        const int k0 = tk * 8;
        if (k0 >= N) break;
        sweep_step<0, FWD>(A, ti, tj, tk, k0 + 0, N, T, Np, cbuf, pbuf, piv, active, gid, nthreads); if (k0 + 1 >= N) break;
        sweep_step<1, FWD>(A, ti, tj, tk, k0 + 1, N, T, Np, cbuf, pbuf, piv, active, gid, nthreads); if (k0 + 2 >= N) break;
        sweep_step<2, FWD>(A, ti, tj, tk, k0 + 2, N, T, Np, cbuf, pbuf, piv, active, gid, nthreads); if (k0 + 3 >= N) break;
        sweep_step<3, FWD>(A, ti, tj, tk, k0 + 3, N, T, Np, cbuf, pbuf, piv, active, gid, nthreads); if (k0 + 4 >= N) break;
        sweep_step<4, FWD>(A, ti, tj, tk, k0 + 4, N, T, Np, cbuf, pbuf, piv, active, gid, nthreads); if (k0 + 5 >= N) break;
        sweep_step<5, FWD>(A, ti, tj, tk, k0 + 5, N, T, Np, cbuf, pbuf, piv, active, gid, nthreads); if (k0 + 6 >= N) break;
        sweep_step<6, FWD>(A, ti, tj, tk, k0 + 6, N, T, Np, cbuf, pbuf, piv, active, gid, nthreads); if (k0 + 7 >= N) break;
        sweep_step<7, FWD>(A, ti, tj, tk, k0 + 7, N, T, Np, cbuf, pbuf, piv, active, gid, nthreads);
    }
    }

    // ---- log-determinant, info, quadratic form ---------------------------------------------------
    const int tN = N >> 3, rN = N & 7;
    double ld = 0.0;
    int bad = INT_MAX;
    for (int k = tid; k < N; k += nthreads) {
        const double d = piv[k];
        if (!(d > 0.0)) bad = min(bad, k + 1); else ld += log(d);
    }
    ld = block_sum(ld, red, tid, nthreads, gid);
    if (tid == 0) *s_bad_p = INT_MAX;
    GROUP_SYNC();
    if (bad != INT_MAX) atomicMin(s_bad_p, bad);   // min is order independent: deterministic
    GROUP_SYNC();
    const int info = (*s_bad_p == INT_MAX) ? 0 : *s_bad_p;

    if (active && ti == tN && tj == tN) {
        double qv = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r == rN) qv = -A[r][r];
        red[32] = qv;
    }
    GROUP_SYNC();
    const double quad = red[32];
    const double ll = -0.5 * ((double)N * LOG2PI + ld + quad);   // logpdf(MvNormal(bbar,K), Y)  (:139)
    if (tid == 0) {
        b.ll[e] = info ? -INFINITY : ll;
        if (b.info) b.info[e] = info;
    }
    if (FWD || !b.want_grad) return;
    if (info) {
        if (tid <= L) b.grad[(size_t)e * (L + 1) + tid] = 0.0;
        return;
    }

    // ---- gradient: W = a a' - K~^-1 contracted with K and dK/drho ----------------------------------
    if (active && ti == tN) {   // border row holds a = K~^-1 r
        double vals[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double x = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) if (r == rN) x = A[r][c];
            vals[c] = (tj * 8 + c < N) ? x : 0.0;
        }
        store8(abuf, tj, T, vals);
    }
    GROUP_SYNC();

    if (b.dump_kinv && active) {   // K~^-1 = -(swept matrix); written once per gpcc call (postb / pred), not in the fit loop
        double* out = b.dump_kinv + (size_t)e * N * N;
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int i = ti * 8 + r, j = tj * 8 + c;
                if (i < N && j < N && j <= i) { out[(size_t)j * N + i] = -A[r][c]; out[(size_t)i * N + j] = -A[r][c]; }
            }
    }
    if (b.dump_a) for (int i = tid; i < N; i += nthreads) b.dump_a[(size_t)e * N + i] = abuf[cidx(i, T)];

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
            const int cj = cidx(tj * 8 + half * 4 + cc, T);
            tc[cc] = tsh[cj]; ac[cc] = av[cj]; wc[cc] = abuf[cj]; cols[cc] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ci = cidx(ti * 8 + r, T);
            const double tr = tsh[ci], ar = av[ci], wr = abuf[ci];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = half * 4 + cc;
                const double W = fma(wr, wc[cc], A[r][c]);       // a_i a_j - (K~^-1)_ij
                double kv, dkv;
                kern_value_drho<KID>(tr - tc[cc], kp, kv, dkv);
                const double aa = ar * ac[cc];                   // 0 on padding / border rows
                double ct = W * (aa * kv);
                double et = W * (aa * dkv);
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
                // fold the column sums of the strictly-lower part into the same slot as the row sums
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
    es = block_sum(active ? es : 0.0, red, tid, nthreads, gid);   // (contains the __syncthreads that orders `part`)

    // s_i = sum_j W_ij K_ij (full row);  dlogL/dalpha_p = (1/alpha_p) sum_{i in band p} s_i
    double* srow = cbuf;   // natural layout, reuse
    for (int i = tid; i < N; i += nthreads) {
        const double* pp = part + (size_t)(i >> 3) * T * 8 + (i & 7);
        double s = 0.0;
        for (int src = 0; src < T; ++src) s += pp[src * 8];
        srow[i] = s;
    }
    GROUP_SYNC();
    const int warp = tid >> 5, lane = tid & 31, nwarps = (nthreads + 31) >> 5;
    for (int pb = warp; pb < L; pb += nwarps) {
        double s = 0.0;
        for (int i = p.band_start[pb] + lane; i < p.band_start[pb + 1]; i += 32) s += srow[i];
        s = warp_sum(s);
        if (lane == 0) b.grad[(size_t)e * (L + 1) + pb] = s / b.alpha[(size_t)e * L + pb];
    }
    if (tid == 0) b.grad[(size_t)e * (L + 1) + L] = es;   // 0.5 * sum_full = sum over the strict lower triangle
}

size_t smem_bytes(int T, int want_grad) {
    const int Np = T * TS;
    size_t doubles = (size_t)VLEN * 10 + 4 + 64 + (want_grad ? (size_t)T * T * 8 : 0);
    return doubles * sizeof(double) + (size_t)Np * sizeof(int) + 16;
}

template <int KID, int MT, int MB, int NMAT, bool BF, bool PAIR, bool FWD>
void go(const DevProblem& p, const EvalBatch& b, int T, int gd, int threads, int maxT, cudaStream_t s) {
    auto kfn = small_sweep_kernel<KID, MT, MB, NMAT, BF, PAIR, FWD>;
    const size_t group_bytes = (size_t)gd * 8;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, NMAT * 8 * (int)((smem_bytes(maxT, 1) + 15) / 16 * 2));
    kfn<<<(b.M + NMAT - 1) / NMAT, NMAT * threads, NMAT * group_bytes, s>>>(p, b, T, gd);
}

template <int KID>
cudaError_t launch_kid(const DevProblem& p, const EvalBatch& b, int T, cudaStream_t s) {
    const int ntiles = T * (T + 1) / 2;
    const int threads = (ntiles + 31) / 32 * 32;
    const size_t sm = smem_bytes(T, b.want_grad);
    static const int variant_env = getenv("GPCC_SMALL_VARIANT") ? atoi(getenv("GPCC_SMALL_VARIANT")) : -1;
    static const bool no_fwd = getenv("GPCC_SMALL_NO_FWD") != nullptr;
    static int nsm = 0;
    if (nsm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    }
    // Default: pick the occupancy variant by batch size (measured on B200, profiles/small_sweep_variants_r1.log).
    //   N <= 111 (128 threads): three CTAs per SM are 8 % faster once every SM has three (32.7 vs 35.6 us per evaluation
    //   and SM), but slower when the batch leaves SMs with one or two (82 vs 65 us at one CTA per SM);
    //   N <= 151 (192 threads): a CTA that is alone on its SM runs faster with 255 registers and no spills (101 vs 122 us),
    //   two co-resident CTAs at 168 registers win as soon as there is more than one evaluation per SM (134 vs 188 us for two).
    int variant = variant_env;
    if (variant < 0) variant = (threads <= 128) ? (b.M >= 3 * nsm ? 2 : 1) : (b.M <= nsm ? 0 : 1);
    const int gd = (int)((sm + 15) / 16 * 2);   // doubles per group, 16-byte aligned
    // forward-only mode (N^3/3 flop): no gradient and nothing that needs the inverse or a = K~^-1 r
    const bool fwd = !no_fwd && !b.want_grad && !b.dump_kinv && !b.dump_a && variant_env < 0;
    if (fwd) {
        if (threads <= 128) {
            if (variant == 2) go<KID, 128, 3, 1, false, false, true>(p, b, T, gd, threads, 15, s);
            else              go<KID, 128, 2, 1, false, false, true>(p, b, T, gd, threads, 15, s);
        } else if (threads <= 192) go<KID, 192, 2, 1, false, false, true>(p, b, T, gd, threads, 19, s);
        else if (threads <= 224)   go<KID, 224, 1, 1, false, false, true>(p, b, T, gd, threads, 20, s);
        else                       go<KID, 352, 1, 1, false, false, true>(p, b, T, gd, threads, SMALL_MAX_T, s);
        return cudaGetLastError();
    }
    if (threads <= 128) {
        if (variant == 3)      go<KID, 384, 1, 3, false, false, false>(p, b, T, gd, threads, 15, s);   // three matrices per CTA: 12 warps, 3 per scheduler
        else if (variant == 2) go<KID, 128, 3, 1, false, false, false>(p, b, T, gd, threads, 15, s);
        else                   go<KID, 128, 2, 1, false, false, false>(p, b, T, gd, threads, 15, s);
    } else if (threads <= 192 && variant == 3) {   // two matrices per CTA: 12 warps, 3 per scheduler
        go<KID, 384, 1, 2, false, false, false>(p, b, T, gd, threads, 19, s);
    } else if (threads <= 192 && variant == 5 && (p.N % 2) == 0) {   // two pivots per barrier
        go<KID, 192, 2, 1, false, true, false>(p, b, T, gd, threads, 19, s);
    } else if (threads <= 192 && variant == 6 && (p.N % 2) == 0) {   // two pivots per barrier, one CTA per SM at 255 registers (no spills)
        go<KID, 192, 1, 1, false, true, false>(p, b, T, gd, threads, 19, s);
    } else if (threads <= 192 && variant == 4) {   // branch-free step
        go<KID, 192, 2, 1, true, false, false>(p, b, T, gd, threads, 19, s);
    } else if (threads <= 192 && variant == 1) {
        go<KID, 192, 2, 1, false, false, false>(p, b, T, gd, threads, 19, s);
    } else if (threads <= 224) {
        go<KID, 224, 1, 1, false, false, false>(p, b, T, gd, threads, 20, s);
    } else {
        go<KID, 352, 1, 1, false, false, false>(p, b, T, gd, threads, SMALL_MAX_T, s);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t small_sweep_init() { return cudaSuccess; }

cudaError_t small_sweep_launch(const DevProblem& p, const EvalBatch& b, cudaStream_t s) {
    const int T = (p.N + 1 + TS - 1) / TS;
    switch (p.kernel_id) {
        case K_OU:  return launch_kid<K_OU>(p, b, T, s);
        case K_RBF: return launch_kid<K_RBF>(p, b, T, s);
        case K_M32: return launch_kid<K_M32>(p, b, T, s);
        case K_M52: return launch_kid<K_M52>(p, b, T, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace gpcc
