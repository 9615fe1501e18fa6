"""Scalar GP kernels of the reference (oracle; test infrastructure only).

Follows /root/reference/src/util.jl:15-52.  All kernels have k(0)=1 and take the lengthscale rho
*linearly* where the reference does (note rbf = exp(-0.5*d^2/(2*rho)) = exp(-d^2/(4 rho)), util.jl:28).
The d/drho derivatives are not in the reference (it has no gradient code); they are needed by the
analytic hyper-parameter gradient that north_star adds and are checked against finite differences.
"""
import numpy as np

SQRT3 = np.sqrt(3.0)
SQRT5 = np.sqrt(5.0)

# enum shared with include/gpcc_b200.h
KERNEL_IDS = {"OU": 0, "rbf": 1, "matern32": 2, "matern52": 3}
KERNELS = tuple(KERNEL_IDS)


def kernel_value(name, d, rho):
    """k(xi, xj; rho) as a function of d = xi - xj (array ok)."""
    d = np.asarray(d, dtype=np.float64)
    if name == "OU":                       # util.jl:15-23
        return np.exp(-np.abs(d) / rho)
    if name == "rbf":                      # util.jl:28
        return np.exp(-0.5 * d * d / (2.0 * rho))
    if name == "matern32":                 # util.jl:32-40
        r = np.abs(d)
        return (1.0 + SQRT3 * r / rho) * np.exp(-SQRT3 * r / rho)
    if name == "matern52":                 # util.jl:44-52
        r = np.abs(d)
        return (1.0 + SQRT5 * r / rho + (5.0 * r * r) / (3.0 * rho * rho)) * np.exp(-SQRT5 * r / rho)
    raise ValueError(f"unknown kernel {name!r}")


def kernel_drho(name, d, rho):
    """d k / d rho."""
    d = np.asarray(d, dtype=np.float64)
    r = np.abs(d)
    if name == "OU":
        return np.exp(-r / rho) * r / (rho * rho)
    if name == "rbf":
        return np.exp(-d * d / (4.0 * rho)) * d * d / (4.0 * rho * rho)
    if name == "matern32":
        a = SQRT3 * r / rho
        return a * a * np.exp(-a) / rho
    if name == "matern52":
        a = SQRT5 * r / rho
        return a * a * (1.0 + a) * np.exp(-a) / (3.0 * rho)
    raise ValueError(f"unknown kernel {name!r}")
