// Scalar GP kernels on the device (FP64).  Restates /root/reference/src/util.jl:15-52; the d/drho forms
// are SURVEY.md section 8 row a1 (the reference has no gradient code).  `KernParams` hoists the
// per-evaluation divisions by rho out of the N^2 element loop (one rounding apart from the reference's
// operation order, i.e. ~1e-16 relative per element).
#pragma once
#include <cuda_runtime.h>

namespace gpcc {

enum { K_OU = 0, K_RBF = 1, K_M32 = 2, K_M52 = 3 };

struct KernParams {
    double c1;       // OU: 1/rho   rbf: 1/(4 rho)   m32: sqrt3/rho   m52: sqrt5/rho
    double inv_rho;  // 1/rho
};

__host__ __device__ __forceinline__ KernParams make_kern_params(int kid, double rho) {
    KernParams p;
    p.inv_rho = 1.0 / rho;
    switch (kid) {
        case K_OU:  p.c1 = 1.0 / rho; break;
        case K_RBF: p.c1 = 0.25 / rho; break;
        case K_M32: p.c1 = 1.7320508075688772935 / rho; break;
        default:    p.c1 = 2.2360679774997896964 / rho; break;
    }
    return p;
}

// k(d; rho) with d = x_i - x_j
template <int KID>
__device__ __forceinline__ double kern_value(double d, const KernParams& p) {
    if (KID == K_OU) {                       // util.jl:15-23   exp(-r/rho)
        return exp(-fabs(d) * p.c1);
    } else if (KID == K_RBF) {               // util.jl:28      exp(-0.5 d^2 / (2 rho))
        return exp(-(d * d) * p.c1);
    } else if (KID == K_M32) {               // util.jl:32-40   (1+a) exp(-a), a = sqrt3 r / rho
        const double a = fabs(d) * p.c1;
        return (1.0 + a) * exp(-a);
    } else {                                 // util.jl:44-52   (1 + a + a^2/3) exp(-a), a = sqrt5 r / rho
        const double a = fabs(d) * p.c1;
        return (1.0 + a + (a * a) * (1.0 / 3.0)) * exp(-a);
    }
}

// k and dk/drho in one pass (one exp)
template <int KID>
__device__ __forceinline__ void kern_value_drho(double d, const KernParams& p, double& k, double& dk) {
    if (KID == K_OU) {                       // dk = k r / rho^2
        const double a = fabs(d) * p.c1;
        k = exp(-a);
        dk = k * a * p.inv_rho;
    } else if (KID == K_RBF) {               // dk = k d^2 / (4 rho^2)
        const double a = (d * d) * p.c1;
        k = exp(-a);
        dk = k * a * p.inv_rho;
    } else if (KID == K_M32) {               // dk = a^2 e^-a / rho
        const double a = fabs(d) * p.c1;
        const double e = exp(-a);
        k = (1.0 + a) * e;
        dk = (a * a) * e * p.inv_rho;
    } else {                                 // dk = a^2 (1+a) e^-a / (3 rho)
        const double a = fabs(d) * p.c1;
        const double e = exp(-a);
        k = (1.0 + a + (a * a) * (1.0 / 3.0)) * e;
        dk = (a * a) * (1.0 + a) * e * ((1.0 / 3.0) * p.inv_rho);
    }
}

}  // namespace gpcc
