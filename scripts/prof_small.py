"""Profiling driver: a few launches of the fused small-N kernel (N=150, matern32, logL+grad)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import oracle
from gpcc_b200 import Problem, Context
ctx = Context(1)
t, y, s, d = oracle.simulatethreelightcurves()
p = Problem(t, y, s, "matern32", ctx)
M = 148 * 8
rg = np.random.default_rng(1)
delays = np.zeros((M, 3)); delays[:, 1:] = rg.uniform(0, 20, (M, 2))
alpha = rg.uniform(0.5, 3.0, (M, 3)); rho = rg.uniform(0.5, 20, M)
for it in range(3):
    ll, g, info = p.loglik_batch(delays, alpha, rho, want_grad=True)
print("ok", ll[:3], info.max())
