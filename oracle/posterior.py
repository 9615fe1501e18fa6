"""Posterior over candidate delays and the delay prior (oracle; test infrastructure only).

Follows /root/reference/src/getprobabilities.jl:1-20 and /root/reference/src/uniformpriordelay.jl:10-16.
"""
import numpy as np
from scipy.special import logsumexp


def getprobabilities(loglikel, logpriorpdfvalues=None):
    loglikel = np.asarray(loglikel, dtype=np.float64)
    if logpriorpdfvalues is None:
        logpriorpdfvalues = np.ones(loglikel.shape)            # getprobabilities.jl:3 ("flat prior" of ones)
    joint = loglikel + np.asarray(logpriorpdfvalues, dtype=np.float64)   # :14
    return np.exp(joint - logsumexp(joint))                    # :16 (shape preserving)


class Uniform:
    """Minimal stand-in for Distributions.Uniform(a, b): logpdf is -log(b-a) on [a,b], -Inf outside."""

    def __init__(self, a, b):
        self.a, self.b = float(a), float(b)

    def logpdf(self, x):
        x = np.asarray(x, dtype=np.float64)
        inside = (x >= self.a) & (x <= self.b)
        return np.where(inside, -np.log(self.b - self.a), -np.inf)

    def __repr__(self):
        return f"Uniform(a={self.a}, b={self.b})"


def uniformpriordelay(*, L, z):
    upper = 10.0 ** 1.559 * (L * 10.0 ** (-44)) ** 0.549 * (1 + z)   # uniformpriordelay.jl:12
    return Uniform(0.0, upper)                                        # :14
