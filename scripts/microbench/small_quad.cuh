// EXPERIMENT (round 2, measured slower, not part of the library): quad-layout form of the fused small-N evaluator
// (N + 1 <= 152).  Same arithmetic as gpcc_b200/csrc/small_eval.cuh, element by element and in the same order: logL agrees
// with eval_one to 0-2 ulp (58 of 2368 evaluations differ in the last bit through FMA contraction in the assembly), the
// gradient to 4e-13.  Driver: quad_time.cu.
//
// Idea.  Four warps per matrix instead of six: the lower triangle is cut into quads of 4 rows x 8 columns, a thread owns
// THREE quads of one column block (96 doubles = 192 of its 255 registers), N = 150 needs 380 quads = 127 threads, two CTAs
// per SM.  Per step a thread reads 8 column values and 3 x 4 row values for 96 DFMAs (20 broadcast values per 96 DFMAs
// instead of 16 per 64); the quads that do not divide into threes (all in the last block row) go pair + single into a few
// "mixed" threads whose third quad belongs to another column block, hence the second set of column values for slot 2.
// Per-point vectors are laid out [i mod 8][i div 8] so that every access is a 64-bit LDS/STS: with 192 accumulators live
// ptxas has no free aligned register quads, lands every 128-bit load in one scratch quad and copies it out (4 moves each).
//
// Outcome on B200 (profiles/README.md, "round 2, last session"): 81.9 us per evaluation and SM against 60.8 for the
// tile-per-thread kernel (forward-only 63.0 against 42.2).  Why: (1) the premise was wrong -- the hardware already spreads
// the 12 warps of two 6-warp CTAs 3/3/3/3 over the four schedulers (sched_map.cu), there was no 4/4/2/2 imbalance to
// remove; (2) a warp issues in order and a step costs it the sum of the stall counts of its instructions (DFMA 2, STS 4,
// ISETP->BRA 13, ...: ~750 cycles for the ~300 instructions of a quad step, ~480 for the ~190 of a tile step) plus barrier
// and load latency; with two warps per scheduler instead of three nothing hides that chain (issue slots 25 % used, FP64
// pipe 24 %); (3) 96 inlined kernel evaluations per thread make assembly and gradient 85 + 96 KB of straight-line code
// that two CTAs in different phases fetch past the instruction cache (stall_no_inst 48 % / 75 % of those phases).
#pragma once
#include "small_eval.cuh"   // -I gpcc_b200/csrc

namespace gpcc {
namespace small {

constexpr int QMAX_T = 19;            // N + 1 <= 152
constexpr int QS8 = 20;               // doubles per part (>= QMAX_T + 1: one dummy column block behind the data)
constexpr int QVL = 8 * QS8;          // doubles per per-point vector
constexpr int QPIV = 160;             // pivots, natural order
constexpr int QCM = 14;               // threads per column block that can hold column partials (<= 2*QMAX_T/3 + 1)

// per-point vectors: [i mod 8][i div 8].  No two values that one thread reads or writes are adjacent, on purpose: every
// access stays a 64-bit LDS/STS that lands in any aligned register pair.  With 192 accumulator registers live ptxas finds no
// free aligned register quads; it lands every 128-bit load (which it also builds by itself from adjacent 64-bit ones) in one
// scratch quad and copies it out again, four moves per load, and gathers every 128-bit store the same way.
__device__ __forceinline__ int qidx(int i) { return (i & 7) * QS8 + (i >> 3); }
__device__ __forceinline__ int qrowoff(int rq) { return 4 * (rq & 1) * QS8 + (rq >> 1); }   // + r * QS8 for row r of the quad

__device__ __forceinline__ void qloadcol(const double* buf, int tj, double (&v)[8]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = buf[c * QS8 + tj];
}
__device__ __forceinline__ void qloadrow(const double* buf, int ro, double (&x)[4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) x[r] = buf[ro + r * QS8];
}

// threads the layout needs for T block rows
__host__ __device__ inline int quad_threads(int T) {
    int n = 0;
    for (int c = 0; c < T; ++c) { const int q = 2 * (T - c); n += q / 3 + (q % 3 == 2 ? 1 : 0); }
    return n;
}
__host__ __device__ inline size_t qeval_smem_bytes(int T, int want_grad) {
    const size_t doubles = (size_t)QVL * 7 + QPIV + 4 + 64 + (want_grad ? (size_t)2 * T * T * 4 + (size_t)T * QCM * 8 : 0);
    return doubles * sizeof(double) + (size_t)(8 * T + 8) * sizeof(int) + 16;
}

// Which quads does thread `tid` own?  Column block c holds the row quads 2c .. 2T-1.  Regular threads take them three at a
// time from the top (thread m of block c: quads 2c+3m ..+2; the first thread of a block holds the diagonal tile in slots 0
// and 1); what is left over -- a pair (2T-2, 2T-1) or a single (2T-1) -- goes pair + single into the mixed threads behind the
// regular ones.  Empty slots point at the dummy block (row quad 2T, column block T: zeros, never owners of anything).
struct QuadOwner {
    int tja, tjb;        // column block of slots 0,1 / of slot 2
    int rq[3];           // row quad of each slot
    int ma, mb;          // index of this thread among the holders of column block tja / tjb (column partials of the gradient)
};
__device__ __forceinline__ QuadOwner quad_layout(int T, int tid) {
    QuadOwner q;
    q.tja = q.tjb = T;
    q.rq[0] = q.rq[1] = q.rq[2] = 2 * T;
    q.ma = q.mb = 0;
    int base = 0;
    bool found = false;
    for (int c = 0; c < T; ++c) {
        const int full = (2 * (T - c)) / 3;
        if (!found && tid < base + full) {
            const int m = tid - base;
            q.tja = q.tjb = c;
            q.rq[0] = 2 * c + 3 * m; q.rq[1] = q.rq[0] + 1; q.rq[2] = q.rq[0] + 2;
            q.ma = q.mb = m;
            found = true;
        }
        base += full;
    }
    if (!found) {
        const int i = tid - base;
        int ip = 0, is = 0;
        for (int c = 0; c < T; ++c) {
            const int n = 2 * (T - c), full = n / 3, rem = n - 3 * full;
            if (rem == 2) { if (ip == i) { q.tja = c; q.rq[0] = 2 * T - 2; q.rq[1] = 2 * T - 1; q.ma = full; } ++ip; }
            if (rem == 1) { if (is == i) { q.tjb = c; q.rq[2] = 2 * T - 1; q.mb = full; } ++is; }
        }
    }
    return q;
}

// Publish column kn = 8 tkn + KKN of the symmetric matrix for the next step (chunk layout, 16-byte stores).
//   column part (rows > kn): quads of column block tkn below the pivot quad, column KKN of the quad;
//   diagonal quad (row quad qkn of column block tkn: always slot KKN>>2 of the first thread of the block): rows >= kn from the
//     column, rows < kn of the pivot tile from the pivot row; the pivot and its reciprocal;
//   row part (columns < 8 tkn; not read in forward-only mode): quads of row quad qkn, row KKN&3 of the quad.
template <int KKN>
__device__ __forceinline__ void qpublish(const double (&A)[3][4][8], const QuadOwner& q, int tkn, double* nb, double* pslot,
                                         double* pivslot, double prn, bool fwd) {
    constexpr int RN = KKN & 3, SN = KKN >> 2;
    const int qkn = 2 * tkn + SN;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int tj = s < 2 ? q.tja : q.tjb;
        const int rq = q.rq[s];
        if (tj == tkn) {
            if (rq > qkn) {
                double* d = nb + qrowoff(rq);
#pragma unroll
                for (int r = 0; r < 4; ++r) d[r * QS8] = A[s][r][KKN];
            } else if (s == SN && rq == qkn) {
                double* d = nb + tkn;
#pragma unroll
                for (int e = 0; e < 4 * SN + 4; ++e)   // rows of the pivot tile above the pivot lie in the pivot row
                    d[e * QS8] = (e < KKN) ? A[s][RN][e] : A[s][e & 3][KKN];
                *pivslot = A[s][RN][KKN];
                *pslot = prn;
            }
        } else if (!fwd && rq == qkn) {
            double* d = nb + tj;
#pragma unroll
            for (int c = 0; c < 8; ++c) d[c * QS8] = A[s][RN][c];
        }
    }
}

// One quad of one sweep step: A_ij -= x_i (v_j / d).  `v` is already scaled by 1/d and carries the column-owner multiplier.
template <int KK, int S_>
__device__ __forceinline__ void qupdate(double (&A)[3][4][8], const double* cb, int ro, bool own_row, const double (&v)[8]) {
    constexpr int R = KK & 3;
    double x[4];
    qloadrow(cb, ro, x);
    if (own_row) x[R] -= 1.0;               // row k of its owners becomes c/d (see sweep_step in small_eval.cuh)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) A[S_][r][c] = fma(-x[r], v[c], A[S_][r][c]);
}

template <int KK>
__device__ __forceinline__ void qstep(double (&A)[3][4][8], const QuadOwner& q, int tk, int k, double* cbuf, double* pbuf, double* piv,
                                      bool fwd) {
    constexpr int PAR = KK & 1, NPAR = PAR ^ 1;      // k = 8 tk + KK: the buffer parity is known at compile time
    constexpr int KKN = (KK + 1) & 7;
    constexpr int R = KK & 3, S = KK >> 2;           // pivot row inside its quad; slot of the diagonal quad in its owner
    constexpr int RN = KKN & 3, SN = KKN >> 2;       // the same for the next pivot
    const int qk = 2 * tk + S;
    const double* cb = cbuf + PAR * QVL;
    const double pr = pbuf[PAR];
    const bool liveA = !fwd || q.tja >= tk, liveB = !fwd || q.tjb >= tk;
    double prn = 0.0;
    if (liveA) {
        double v[8];
        qloadcol(cb, q.tja, v);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] *= pr;
        const bool own_col = q.tja == tk;
        if (own_col) v[KK] = 1.0 - pr;               // column k of its owners becomes c/d
        // the slot that may hold the next pivot first: its reciprocal is formed by every thread right behind it, branch free,
        // and its latency hides under the remaining 64 DFMAs
        qupdate<KK, SN>(A, cb, qrowoff(q.rq[SN]), q.rq[SN] == qk, v);
        if (SN == S && own_col && q.rq[S] == qk) A[S][R][KK] = -pr;
        prn = fast_rcp(A[SN][RN][KKN]);
        qupdate<KK, 1 - SN>(A, cb, qrowoff(q.rq[1 - SN]), q.rq[1 - SN] == qk, v);
        if (SN != S && own_col && q.rq[S] == qk) A[S][R][KK] = -pr;
    }
    if (liveB) {
        double v[8];
        qloadcol(cb, q.tjb, v);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] *= pr;
        if (q.tjb == tk) v[KK] = 1.0 - pr;
        qupdate<KK, 2>(A, cb, qrowoff(q.rq[2]), q.rq[2] == qk, v);
    }
    qpublish<KKN>(A, q, KK < 7 ? tk : tk + 1, cbuf + NPAR * QVL, pbuf + NPAR, piv + k + 1, prn, fwd);
    __syncthreads();
}

// One evaluation by the whole CTA (quad layout).  Same contract as eval_one; T <= QMAX_T, blockDim.x >= quad_threads(T).
// `smem`: qeval_smem_bytes(T, want_grad) bytes, 16-byte aligned.
template <int KID>
__device__ __forceinline__ void eval_one_q(const DevProblem& p, int T, const QuadOwner& q, double* smem, const double* delays_e,
                                           const double* alpha_e, double rho, bool want_grad, bool fwd, double* out_ll,
                                           double* out_grad, int* out_info) {
    const int N = p.N, L = p.L;
    const int nthreads = blockDim.x;
    const int tid = threadIdx.x;

    double* tsh = smem;             // shifted times t_i - tau_band(i)
    double* av = tsh + QVL;         // alpha_band(i), 0 for padding
    double* sbv = av + QVL;         // Sigma_b[band(i)]
    double* dadd = sbv + QVL;       // sigma_i^2
    double* abuf = dadd + QVL;      // residual r, later a = K~^-1 r
    double* cbuf = abuf + QVL;      // 2 x broadcast column
    double* piv = cbuf + 2 * QVL;   // pivots (natural)
    double* pbuf = piv + QPIV;      // 2 pivot reciprocals (+2 pad)
    double* red = pbuf + 4;         // 64 reduction slots
    int* s_bad_p = reinterpret_cast<int*>(red + 48);
    double* rowpart = red + 64;                                  // [2T][T][4] row sums of the quads        (gradient only)
    double* colpart = rowpart + (want_grad ? 2 * T * T * 4 : 0); // [T][QCM][8] column sums per holder thread (gradient only)
    int* bandv = reinterpret_cast<int*>(colpart + (want_grad ? T * QCM * 8 : 0));   // [8T + 8] natural

    const KernParams kp = make_kern_params(KID, rho);

    for (int i = tid; i < 8 * T + 8; i += nthreads) {
        const int ci = qidx(i);
        if (i < N) {
            const int bi = p.band[i];
            tsh[ci] = p.t[i] - delays_e[bi];   // delayedCovariance.jl:27 (x - delays[l])
            av[ci] = alpha_e[bi];
            sbv[ci] = p.sigb[i];
            dadd[ci] = p.s2[i];
            abuf[ci] = p.resid[i];
            bandv[i] = bi;
        } else {
            tsh[ci] = 0.0; av[ci] = 0.0; sbv[ci] = 0.0; dadd[ci] = 0.0; abuf[ci] = 0.0;
            bandv[i] = -1 - i;
            if (i >= 8 * T) { cbuf[ci] = 0.0; cbuf[QVL + ci] = 0.0; }   // the dummy block stays zero
        }
    }
    __syncthreads();

    // ---- assembly of the bordered matrix in registers ------------------------------------------------
    double A[3][4][8];
#pragma unroll
    for (int g = 0; g < 2; ++g) {            // g = 0: slots 0,1 (column block tja); g = 1: slot 2 (tjb)
        const int tj = g ? q.tjb : q.tja;
        double tc[8], ac[8], rc[8];
        int bc[8];
        qloadcol(tsh, tj, tc);
        qloadcol(av, tj, ac);
        qloadcol(abuf, tj, rc);
#pragma unroll
        for (int c = 0; c < 8; ++c) bc[c] = bandv[tj * 8 + c];
#pragma unroll
        for (int ss = 0; ss < 2; ++ss) {
            if (g == 1 && ss == 1) continue;
            const int s = g ? 2 : ss;
            const int rq = q.rq[s];
            double tr[4], ar[4], sbr[4], dr[4];
            qloadrow(tsh, qrowoff(rq), tr);
            qloadrow(av, qrowoff(rq), ar);
            qloadrow(sbv, qrowoff(rq), sbr);
            qloadrow(dadd, qrowoff(rq), dr);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = rq * 4 + r;
                const int br = bandv[i];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = tj * 8 + c;
                    const double kv = kern_value<KID>(tr[r] - tc[c], kp);
                    double val = (ar[r] * ac[c]) * kv;       // scale[l]*scale[m]*kernel  (delayedCovariance.jl:27)
                    if (i == j) val += dr[r];                // + Sobs                   (gpccfixdelay_marginaliseb.jl:135)
                    if (br == bc[c]) val += sbr[r];          // + B = Q Sigma_b Q'
                    if (i == N) val = rc[c];                 // border row: r = Y - bbar (corner = 0)
                    if (i > N && i == j) val = 1.0;          // padding
                    A[s][r][c] = val;
                }
            }
        }
    }
    __syncthreads();   // everyone has read abuf/tsh before cbuf traffic starts (abuf is reused later)

    // ---- publish column 0, then N sweep steps ----------------------------------------------------
    qpublish<0>(A, q, 0, cbuf, pbuf, piv, fast_rcp(A[0][0][0]), fwd);
    __syncthreads();
    for (int tk = 0; 8 * tk < N; ++tk) {   // one copy of the eight specialised steps; the guards only matter in the last tile
        const int k0 = tk * 8;
        qstep<0>(A, q, tk, k0 + 0, cbuf, pbuf, piv, fwd);
        if (k0 + 1 < N) qstep<1>(A, q, tk, k0 + 1, cbuf, pbuf, piv, fwd);
        if (k0 + 2 < N) qstep<2>(A, q, tk, k0 + 2, cbuf, pbuf, piv, fwd);
        if (k0 + 3 < N) qstep<3>(A, q, tk, k0 + 3, cbuf, pbuf, piv, fwd);
        if (k0 + 4 < N) qstep<4>(A, q, tk, k0 + 4, cbuf, pbuf, piv, fwd);
        if (k0 + 5 < N) qstep<5>(A, q, tk, k0 + 5, cbuf, pbuf, piv, fwd);
        if (k0 + 6 < N) qstep<6>(A, q, tk, k0 + 6, cbuf, pbuf, piv, fwd);
        if (k0 + 7 < N) qstep<7>(A, q, tk, k0 + 7, cbuf, pbuf, piv, fwd);
    }

    // ---- log-determinant, info, quadratic form ---------------------------------------------------
    const int rqN = N >> 2, tN = N >> 3, rN = N & 3, cN = N & 7;
    double ld = 0.0;
    int bad = INT_MAX;
    for (int k = tid; k < N; k += nthreads) {
        const double d = piv[k];
        if (!(d > 0.0)) bad = min(bad, k + 1); else ld += log(d);
    }
    ld = block_sum(ld, red, tid, nthreads);
    if (tid == 0) *s_bad_p = INT_MAX;
    __syncthreads();
    if (bad != INT_MAX) atomicMin(s_bad_p, bad);   // min is order independent: deterministic
    __syncthreads();
    const int info = (*s_bad_p == INT_MAX) ? 0 : *s_bad_p;

#pragma unroll
    for (int s = 0; s < 2; ++s) {                  // the corner lies in the diagonal tile: slot 0 or 1 of its holder
        if (q.rq[s] == rqN && q.tja == tN) {
            double qv = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) if (r == rN && c == cN) qv = -A[s][r][c];
            red[32] = qv;
        }
    }
    __syncthreads();
    const double quad = red[32];
    const double ll = -0.5 * ((double)N * LOG2PI + ld + quad);   // logpdf(MvNormal(bbar,K), Y)  (:139)
    if (tid == 0) {
        *out_ll = info ? -INFINITY : ll;
        *out_info = info;
    }
    if (!want_grad) return;
    if (info) {
        if (tid <= L) out_grad[tid] = 0.0;
        return;
    }

    // ---- gradient: W = a a' - K~^-1 contracted with K and dK/drho (full sweep only: fwd is false here) -----------
#pragma unroll
    for (int s = 0; s < 3; ++s) {                  // border row holds a = K~^-1 r
        const int tj = s < 2 ? q.tja : q.tjb;
        if (q.rq[s] == rqN) {
            double vals[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double x = 0.0;
#pragma unroll
                for (int r = 0; r < 4; ++r) if (r == rN) x = A[s][r][c];
                vals[c] = (tj * 8 + c < N) ? x : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) abuf[c * QS8 + tj] = vals[c];
        }
    }
    __syncthreads();

    double es = 0.0;
    double colsA[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) colsA[c] = 0.0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int tj = s < 2 ? q.tja : q.tjb;
        const int rq = q.rq[s];
        const int dlt = 4 * rq - 8 * tj;           // i - j = dlt + r - c; only the diagonal tile has dlt < 8
        double tr[4], ar[4], wr[4], rows[4];
        qloadrow(tsh, qrowoff(rq), tr);
        qloadrow(av, qrowoff(rq), ar);
        qloadrow(abuf, qrowoff(rq), wr);
#pragma unroll
        for (int r = 0; r < 4; ++r) rows[r] = 0.0;
        double colsB[8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {     // four columns at a time: limits the live column data
            double tc[4], ac[4], wc[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int cj = qidx(tj * 8 + half * 4 + cc);
                tc[cc] = tsh[cj]; ac[cc] = av[cj]; wc[cc] = abuf[cj];
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = half * 4 + cc;
                double csum = 0.0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const double W = fma(wr[r], wc[cc], A[s][r][c]);     // a_i a_j - (K~^-1)_ij
                    double kv, dkv;
                    kern_value_drho<KID>(tr[r] - tc[cc], kp, kv, dkv);
                    const double aa = ar[r] * ac[cc];                    // 0 on padding / border rows
                    double ct = W * (aa * kv);
                    double et = W * (aa * dkv);
                    const int off = dlt + r - c;
                    if (off == 0) { rows[r] += ct; ct = 0.0; et = 0.0; }   // diagonal counted once, dk(0)=0
                    else if (off < 0) { ct = 0.0; et = 0.0; }              // upper part of the diagonal tile is unused
                    rows[r] += ct;
                    csum += ct;
                    es += et;
                }
                if (s < 2) colsA[c] += csum; else colsB[c] = csum;
            }
        }
        if (rq < 2 * T) {
#pragma unroll
            for (int r = 0; r < 4; ++r) rowpart[(rq * T + tj) * 4 + r] = rows[r];
        }
        if (s == 2) {
            if (q.tjb == q.tja) {
#pragma unroll
                for (int c = 0; c < 8; ++c) colsA[c] += colsB[c];
            } else if (q.tjb < T) {
#pragma unroll
                for (int c = 0; c < 8; ++c) colpart[(q.tjb * QCM + q.mb) * 8 + c] = colsB[c];
            }
        }
    }
    if (q.tja < T) {
#pragma unroll
        for (int c = 0; c < 8; ++c) colpart[(q.tja * QCM + q.ma) * 8 + c] = colsA[c];
    }
    es = block_sum(es, red, tid, nthreads);   // (contains the __syncthreads that orders rowpart / colpart)

    // s_i = sum_j W_ij K_ij (full row) = row sums of the quads of row i + column sums of column i below the diagonal;
    // dlogL/dalpha_p = (1/alpha_p) sum_{i in band p} s_i
    double* srow = cbuf;   // natural layout, reuse
    for (int i = tid; i < N; i += nthreads) {
        const int ti = i >> 3;
        const double* rp = rowpart + (size_t)(i >> 2) * T * 4 + (i & 3);
        double s = 0.0;
        for (int tj = 0; tj <= ti; ++tj) s += rp[tj * 4];
        const int nq = 2 * (T - ti), holders = nq / 3 + (nq % 3 ? 1 : 0);
        const double* cp = colpart + (size_t)ti * QCM * 8 + (i & 7);
        for (int m = 0; m < holders; ++m) s += cp[m * 8];
        srow[i] = s;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = (nthreads + 31) >> 5;
    for (int pb = warp; pb < L; pb += nwarps) {
        double s = 0.0;
        for (int i = p.band_start[pb] + lane; i < p.band_start[pb + 1]; i += 32) s += srow[i];
        s = warp_sum(s);
        if (lane == 0) out_grad[pb] = s / alpha_e[pb];
    }
    if (tid == 0) out_grad[L] = es;   // 0.5 * sum_full = sum over the strict lower triangle
}

}  // namespace small
}  // namespace gpcc
