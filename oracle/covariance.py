"""Delayed multi-band block covariance (oracle; test infrastructure only).

Follows /root/reference/src/delayedCovariance.jl:1-38: block (l,m) entry (i,j) is
scale[l]*scale[m]*kernel(x[l][i]-delays[l], y[m][j]-delays[m]; rho)  (line 27), bands concatenated in
order; asserts scale>0 (line 3) and raises for rho<=0 (lines 5-7).  Evaluation order is kept:
subtract each delay from its own time first, then difference, then kernel, then the two scales.
"""
import numpy as np
from .kernels import kernel_value, kernel_drho


def _cat_shifted(x, delays):
    return np.concatenate([np.asarray(xl, dtype=np.float64) - float(d) for xl, d in zip(x, delays)])


def _cat_scale(x, scale):
    return np.concatenate([np.full(len(xl), float(s)) for xl, s in zip(x, scale)])


def delayed_covariance(kernel, scale, delays, rho, x, y=None, drho=False):
    scale = np.asarray(scale, dtype=np.float64)
    if not np.all(scale > 0):
        raise AssertionError("all(scale .> 0)")          # delayedCovariance.jl:3
    if rho <= 0:
        raise ValueError("rho=%.8f is <= 0" % rho)         # delayedCovariance.jl:5-7
    if y is None:
        y = x                                              # delayedCovariance.jl:38
    if not (len(scale) == len(x) == len(y) == len(delays)):
        raise AssertionError("L == length(x) == length(y)")
    tx, ty = _cat_shifted(x, delays), _cat_shifted(y, delays)
    sx, sy = _cat_scale(x, scale), _cat_scale(y, scale)
    d = tx[:, None] - ty[None, :]
    k = kernel_drho(kernel, d, rho) if drho else kernel_value(kernel, d, rho)
    return (sx[:, None] * sy[None, :]) * k
