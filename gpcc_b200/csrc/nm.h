// Nelder-Mead as a state machine (host/device), the reference's own optimiser:
//   optimize(safenegativeobj, theta0, NelderMead(), Optim.Options(iterations, g_tol = 1e-6))   (gpccfixdelay_marginaliseb.jl:205-211)
// Optim.jl is not vendored under /root/reference; this follows its published algorithm the way oracle/fit.py restates it:
// adaptive parameters (Gao & Han 2012: alpha = 1, beta = 1 + 2/n, gamma = 0.75 - 1/(2n), delta = 1 - 1/n), initial simplex
// AffineSimplexer(a = 0.025, b = 0.5), stop when sqrt(var(f_simplex) * n / (n + 1)) < g_tol or after `iterations` iterations.
// It exists for basin-for-basin comparisons with the reference; the default optimiser is the L-BFGS
// of lbfgs.h, which needs ~12x fewer evaluations.  Every evaluation is forward-only (no gradient): N^3/3 flop.
// Driver: small_fit.cu (one persistent CTA per candidate; thread 0 runs this between two evaluations).
#pragma once
#include "lbfgs.h"

namespace gpcc {

constexpr double NM_PENALTY = 1e300;    // safewrapper (:153): a failed evaluation (matrix not PD) counts as a huge objective

struct NmState {
    enum Phase { INIT = 0, REFLECT = 1, EXPAND = 2, CONTRACT_OUT = 3, CONTRACT_IN = 4, SHRINK = 5, DONE = 6 };
    int n = 0, phase = INIT, idx = 0;
    int iters = 0, nfev = 0, status = LbfgsState::RUNNING;
    double x[LBFGS_MAXN + 1][LBFGS_MAXN];   // simplex, kept sorted by f after every iteration
    double f[LBFGS_MAXN + 1];
    double cen[LBFGS_MAXN], xr[LBFGS_MAXN], xt[LBFGS_MAXN];
    double fr = 0.0;
    double fbest = 0.0, xbest[LBFGS_MAXN];   // result

    // insertion sort of the n+1 vertices by f (stable, like the oracle's argsort(kind="stable"))
    GPCC_HD void sort_simplex() {
        for (int i = 1; i <= n; ++i) {
            double fi = f[i], xi[LBFGS_MAXN];
            lb_copy(xi, x[i], n);
            int j = i - 1;
            while (j >= 0 && f[j] > fi) { f[j + 1] = f[j]; lb_copy(x[j + 1], x[j], n); --j; }
            f[j + 1] = fi;
            lb_copy(x[j + 1], xi, n);
        }
    }
    GPCC_HD void finish(int st) {
        int b = 0;
        for (int i = 1; i <= n; ++i) if (f[i] < f[b]) b = i;
        fbest = f[b];
        lb_copy(xbest, x[b], n);
        status = st;
        phase = DONE;
    }
    // after a completed iteration (or the initial simplex): sort, test convergence, set up the reflection point in xt
    GPCC_HD void begin_iteration(int max_iter, double g_tol) {
        sort_simplex();
        double mean = 0.0;
        for (int i = 0; i <= n; ++i) mean += f[i];
        mean /= (n + 1);
        double var = 0.0;
        for (int i = 0; i <= n; ++i) var += (f[i] - mean) * (f[i] - mean);
        var /= (n + 1);                                           // population variance, as numpy's var in oracle/fit.py
        if (sqrt(var * n / (n + 1)) < g_tol) { finish(LbfgsState::CONVERGED); return; }
        if (iters >= max_iter) { finish(LbfgsState::ITER_CAP); return; }
        ++iters;
        for (int k = 0; k < n; ++k) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += x[i][k];
            cen[k] = s / n;
        }
        for (int k = 0; k < n; ++k) { xr[k] = cen[k] + 1.0 * (cen[k] - x[n][k]); xt[k] = xr[k]; }
        phase = REFLECT;
    }
    // x0 with its known objective value f0 (the screening winner); afterwards xt is the first point to evaluate
    GPCC_HD void start(int n_, const double* x0, double f0, int max_iter, double g_tol) {
        n = n_;
        iters = 0; status = LbfgsState::RUNNING;
        nfev = 1;                                                 // f(x0): the value is the screening winner's, not recomputed
        lb_copy(x[0], x0, n);
        f[0] = f0;
        for (int i = 0; i < n; ++i) {                            // AffineSimplexer(a = 0.025, b = 0.5)
            lb_copy(x[i + 1], x0, n);
            x[i + 1][i] = (1.0 + 0.5) * x0[i] + 0.025;
        }
        if (n == 0) { finish(LbfgsState::CONVERGED); return; }
        (void)max_iter; (void)g_tol;
        phase = INIT; idx = 1;
        lb_copy(xt, x[1], n);
    }
    GPCC_HD void replace_worst(const double* xn, double fn) { lb_copy(x[n], xn, n); f[n] = fn; }
    // feed the objective value at xt (ok = false: failed evaluation); afterwards either phase == DONE or xt is the next point
    GPCC_HD void feed(bool ok, double ft, int max_iter, double g_tol) {
        ++nfev;
        if (!ok || !lb_finite(ft)) ft = NM_PENALTY;
        const double beta = 1.0 + 2.0 / n, gamma = 0.75 - 1.0 / (2.0 * n), delta = 1.0 - 1.0 / n;
        switch (phase) {
            case INIT:
                f[idx] = ft;
                if (++idx <= n) { lb_copy(xt, x[idx], n); return; }
                begin_iteration(max_iter, g_tol);
                return;
            case REFLECT:
                fr = ft;
                if (fr < f[0]) {
                    for (int k = 0; k < n; ++k) xt[k] = cen[k] + beta * (xr[k] - cen[k]);
                    phase = EXPAND;
                    return;
                }
                if (fr < f[n - 1]) { replace_worst(xr, fr); begin_iteration(max_iter, g_tol); return; }
                if (fr < f[n]) {
                    for (int k = 0; k < n; ++k) xt[k] = cen[k] + gamma * (xr[k] - cen[k]);       // outside contraction
                    phase = CONTRACT_OUT;
                } else {
                    for (int k = 0; k < n; ++k) xt[k] = cen[k] - gamma * (cen[k] - x[n][k]);     // inside contraction
                    phase = CONTRACT_IN;
                }
                return;
            case EXPAND:
                if (ft < fr) replace_worst(xt, ft); else replace_worst(xr, fr);
                begin_iteration(max_iter, g_tol);
                return;
            case CONTRACT_OUT:
            case CONTRACT_IN: {
                const bool accept = (phase == CONTRACT_OUT) ? (ft <= fr) : (ft < f[n]);
                if (accept) { replace_worst(xt, ft); begin_iteration(max_iter, g_tol); return; }
                for (int i = 1; i <= n; ++i)                                                       // shrink towards the best vertex
                    for (int k = 0; k < n; ++k) x[i][k] = x[0][k] + delta * (x[i][k] - x[0][k]);
                phase = SHRINK; idx = 1;
                lb_copy(xt, x[1], n);
                return;
            }
            case SHRINK:
                f[idx] = ft;
                if (++idx <= n) { lb_copy(xt, x[idx], n); return; }
                begin_iteration(max_iter, g_tol);
                return;
            default:
                return;
        }
    }
};

}  // namespace gpcc
