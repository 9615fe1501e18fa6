"""L-BFGS history length vs evaluations per candidate on the cfg3 grid (dev script)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
from gpcc_b200 import Problem, Context
ctx = Context(1)
t, y, s, d = gpcc_b200.simulatethreelightcurves()
p = Problem(t, y, s, "matern32", ctx)
c = np.arange(0, 20.0001, 0.2)
delays = np.array([[0.0, a, b_] for b_ in c for a in c])
th = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
base = None
for hist in (8, 4, 5, 6, 10, 16):
    for it in range(2):
        t0 = time.time(); r = p.fit_batch(delays, th, iterations=1000, rhomin=0.1, rhomax=300.0, history=hist); dt = time.time() - t0
    if base is None: base = r["loglikel"].copy()
    dl = r["loglikel"] - base
    print("history %2d: %.1f ms, nfev mean %.2f max %d, max loss vs history 8: %.2e, max gain %.2e" % (hist, dt * 1e3, r["nfev"].mean(), r["nfev"].max(), -dl.min(), dl.max()), flush=True)
