"""Profiling driver: three fits of a 2 048-candidate cfg3 sub-grid through the persistent fit kernel (small_fit.cu)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
t, y, s, d = gpcc_b200.simulatethreelightcurves()
p = gpcc_b200.Problem(t, y, s, "matern32")
rg = np.random.default_rng(1)
M = 2048
delays = np.zeros((M, 3)); delays[:, 1:] = rg.uniform(0, 20, (M, 2))
th = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
for it in range(3):
    r = p.fit_batch(delays, th, iterations=1000, rhomin=0.1, rhomax=300.0)
print("ok", r["loglikel"][:3], r["nfev"].mean(), r["nfev"].max())
