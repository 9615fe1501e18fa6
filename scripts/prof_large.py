"""Profiling driver for the tiled large-N path: one wave of 8 matrices, N=6144 (cfg4), matern52."""
import sys, os
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
from gpcc_b200 import Problem, Context
mode = int(os.environ.get("GPCC_PROF_GRAD", "1"))
ctx = Context(1)
t, y, s, d = gpcc_b200.synthetic_bands([2048, 2048, 2048], seed=4)
p = Problem(t, y, s, "matern52", ctx)
M = 8
rg = np.random.default_rng(2)
delays = np.zeros((M, 3)); delays[:, 1:] = rg.uniform(0, 19.8, (M, 2))
alpha = np.tile([1.0, 2.2, 4.0], (M, 1)); rho = np.full(M, 3.5)
out = p.loglik_batch(delays, alpha, rho, want_grad=bool(mode))
print("ok", out[0][:2], out[-1][:2])
