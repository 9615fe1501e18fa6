"""Dev script: per-candidate evaluation counts of the fitted cfg3 grid (for the scheduling study in profiles/README.md)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
t, y, s, d = gpcc_b200.simulatethreelightcurves()
p = gpcc_b200.Problem(t, y, s, "matern32")
c = np.arange(0.0, 20.0001, 0.2)
delays = np.array([[0.0, a, b] for b in c for a in c])
th = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
r = p.fit_batch(delays, th, iterations=1000, rhomin=0.1, rhomax=300.0)
np.savez_compressed("gpurun_out/cfg3_nfev_r2.npz", nfev=r["nfev"], iters=r["iters"], info=r["info"], ll=r["loglikel"])
print("mean", r["nfev"].mean(), "max", r["nfev"].max(), "p50/p90/p99/p99.9", np.percentile(r["nfev"], [50, 90, 99, 99.9]))
