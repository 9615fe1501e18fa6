"""Fit time of the rank-0 shard of the world=W weak-scaling grid, in a single process (dev diagnostic)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
ctx = gpcc_b200.Context(1, profiling=True)
t, y, s, d = gpcc_b200.simulatethreelightcurves()
p = gpcc_b200.Problem(t, y, s, "matern32", ctx)
theta0 = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
for W in (1, 2, 4, 8):
    c2 = np.arange(0.0, 20.0001, 0.2); c3 = np.linspace(0.0, 20.0, 101 * W)
    delays_all = np.array([[0.0, a, b] for b in c3 for a in c2])
    for rank in range(min(W, 2)):
        delays = np.ascontiguousarray(delays_all[rank::W])
        ts = []
        for i in range(4):
            t0 = time.perf_counter(); r = p.fit_batch(delays, theta0, iterations=1000, rhomin=0.1, rhomax=300.0); ts.append((time.perf_counter() - t0) * 1e3)
        st = ctx.stats()
        nf = r["nfev"]
        print("W=%d rank %d: M=%d per-fit ms %s kernel %.0f | nfev mean %.1f max %d, #>100: %d, #>150: %d, #>200: %d status %s" % (
            W, rank, len(delays), ["%.0f" % v for v in ts], st["ms_eval_kernels"], nf.mean(), nf.max(), (nf > 100).sum(), (nf > 150).sum(), (nf > 200).sum(), np.bincount(r["info"] + 1)), flush=True)
