"""Newton-decrement stopping rule for the batched L-BFGS: evaluations saved vs error of the optimum (dev script).
Run once per setting of GPCC_LBFGS_DEC (read once per process); writes loglikel / nfev of the cfg3 grid to gpurun_out/."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
from gpcc_b200 import Problem, Context
dec = os.environ.get("GPCC_LBFGS_DEC", "0")
ctx = Context(1)
t, y, s, d = gpcc_b200.simulatethreelightcurves()
p = Problem(t, y, s, "matern32", ctx)
c = np.arange(0, 20.0001, 0.2)
delays = np.array([[0.0, a, b_] for b_ in c for a in c])
th = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
for it in range(2):
    t0 = time.time(); r = p.grid_posterior(delays, th, iterations=1000, rhomin=0.1, rhomax=300.0); dt = time.time() - t0
print("dec", dec, "time %.1f ms" % (dt * 1e3), "nfev mean %.2f max %d" % (r["nfev"].mean(), r["nfev"].max()), flush=True)
np.savez("gpurun_out/dec_%s.npz" % dec, ll=r["loglikel"], nfev=r["nfev"], post=r["posterior"])
if dec != "0" and os.path.exists("gpurun_out/dec_0.npz"):
    z = np.load("gpurun_out/dec_0.npz")
    dl = r["loglikel"] - z["ll"]
    mass = z["post"] > 1e-12
    print("   vs tight: max |dll| all %.2e, on mass-carrying (%d) %.2e, worst loss %.2e, posterior max diff %.2e" % (
        np.max(np.abs(dl)), mass.sum(), np.max(np.abs(dl[mass])), -dl.min(), np.max(np.abs(r["posterior"] - z["post"]))))
