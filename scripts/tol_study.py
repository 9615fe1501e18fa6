"""How tight must the L-BFGS stopping rule be for |loglikel - optimum| < 1e-6 on the cfg3 grid? (dev script)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
from gpcc_b200 import Problem, Context
ctx = Context(1, profiling=True)
t, y, s, d = gpcc_b200.simulatethreelightcurves()
p = Problem(t, y, s, "matern32", ctx)
c = np.arange(0.0, 20.0001, 0.2)
delays = np.array([[0.0, a, b] for b in c for a in c])
theta0 = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
ref = p.fit_batch(delays, theta0, iterations=1000, rhomin=0.1, rhomax=300.0, gtol=1e-9, ftol=1e-15)
print("ref: mean nfev %.1f max %d iters %.1f status %s" % (ref["nfev"].mean(), ref["nfev"].max(), ref["iters"].mean(), np.bincount(ref["info"] + 1)))
post_ref = gpcc_b200.getprobabilities(ref["loglikel"], ctx=ctx)
for gtol, ftol in [(1e-7, 1e-13)]:
    t0 = time.time()
    r = p.fit_batch(delays, theta0, iterations=1000, rhomin=0.1, rhomax=300.0, gtol=gtol, ftol=ftol)
    dt = time.time() - t0
    gap = ref["loglikel"] - r["loglikel"]
    mass = post_ref > 1e-12
    post = gpcc_b200.getprobabilities(r["loglikel"], ctx=ctx)
    print("gtol %.0e ftol %.0e: %.0f ms, mean nfev %.1f max %d iters %.1f | gap max %.2e (mass-carrying %.2e) n>1e-6: %d, post err %.1e | status %s" % (
        gtol, ftol, dt * 1e3, r["nfev"].mean(), r["nfev"].max(), r["iters"].mean(), gap.max(), gap[mass].max(), (np.abs(gap) > 1e-6).sum(),
        np.abs(post - post_ref).max(), np.bincount(r["info"] + 1)))
