// Host-side check of the L-BFGS state machine that both fit drivers run (gpcc_b200/csrc/lbfgs.h is plain C++ behind GPCC_HD):
// prints "name nfev iters status f x..." per case; tests/test_lbfgs_host.py asserts on the lines.
#include "lbfgs.h"
#include <cstdio>
#include <cmath>
using namespace gpcc;

static double softplus(double x) { return x > 0 ? x + std::log1p(std::exp(-x)) : std::log1p(std::exp(x)); }
static double logistic(double x) { return 1.0 / (1.0 + std::exp(-x)); }

// case 1: Rosenbrock chain in 4 dimensions
static double rosen(const double* x, double* g, int n) {
    double f = 0.0;
    for (int i = 0; i < n; ++i) g[i] = 0.0;
    for (int i = 0; i + 1 < n; ++i) {
        const double a = x[i + 1] - x[i] * x[i], b = 1.0 - x[i];
        f += 100.0 * a * a + b * b;
        g[i] += -400.0 * a * x[i] - 2.0 * b;
        g[i + 1] += 200.0 * a;
    }
    return f;
}
// case 2: the shape of the fit problem: softplus scales with one optimum far up the linear branch (alpha = 150, as for the
// edge-of-grid candidates of the three-band grid) and a logistic length scale
static double scales(const double* th, double* g, int n) {
    const double target[3] = {1.5, 4.0, 150.0};
    double f = 0.0;
    for (int i = 0; i < 3; ++i) {
        const double a = softplus(th[i]), r = std::log(a / target[i]);
        f += r * r;
        g[i] = 2.0 * r / a * logistic(th[i]);
    }
    const double s = logistic(th[3]), rho = 0.1 + 299.9 * s, r = std::log(rho / 25.0);
    f += r * r;
    g[3] = 2.0 * r / rho * 299.9 * s * (1.0 - s);
    (void)n;
    return f;
}

template <class F>
static void run(const char* name, F fun, const double* x0, int n, int n_scale) {
    LbfgsOptions o;
    o.n_scale = n_scale;
    LbfgsState S;
    double g[LBFGS_MAXN];
    const double f0 = fun(x0, g, n);
    S.start(n, x0, f0, g, o);
    while (S.status == LbfgsState::RUNNING) {
        const double ft = fun(S.xt, g, n);
        S.feed(std::isfinite(ft), ft, g, o);
    }
    std::printf("%s %d %d %d %.17g", name, S.nfev, S.iters, S.status, S.f);
    for (int i = 0; i < n; ++i) std::printf(" %.17g", S.x[i]);
    std::printf("\n");
}

int main() {
    const double r0[4] = {-1.2, 1.0, -1.2, 1.0};
    run("rosenbrock", rosen, r0, 4, 0);
    const double s0[4] = {1.0, 2.0, 3.0, -2.0};
    run("scales_fixed_cap", scales, s0, 4, 0);
    run("scales_relative_cap", scales, s0, 4, 3);
    return 0;
}
