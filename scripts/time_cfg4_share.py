"""cfg4 (3 x 2048 points, matern52): the share of rank 0 of the 8-GPU 100 x 100 grid (candidate m -> rank m mod 8: 25 values of tau_2
x 50 values of tau_3), fixed theta, through grid_posterior(iterations=0).  Usage: time_cfg4_share.py [candidates = 1250]; the
A/B switches of the tiled path are environment variables read at first use (GPCC_LARGE_NO_SHARE, GPCC_LARGE_NO_TAUCACHE)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpcc_b200
from gpcc_b200 import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
ctx = gpcc_b200.default_context()
ctx.set_profiling(True)
t4, y4, s4, _ = synthetic.synthetic_bands([2048, 2048, 2048], seed=4)
p4 = gpcc_b200.Problem(t4, y4, s4, gpcc_b200.matern52, ctx)
c4 = np.arange(0.0, 19.8001, 0.2)
grid4 = np.array([[0.0, a, b] for b in c4 for a in c4])
d4 = grid4[np.arange(0, len(grid4), 8)][:n]
th4 = np.concatenate([np.log(np.expm1(np.array([1.0, 2.2, 4.0]))), [np.log((3.5 - 0.1) / (300.0 - 3.5))]])[None]
p4.grid_posterior(d4[:40], th4, iterations=0, rhomin=0.1, rhomax=300.0)
t0 = time.perf_counter()
r = p4.grid_posterior(d4, th4, iterations=0, rhomin=0.1, rhomax=300.0)
dt = time.perf_counter() - t0
st = ctx.stats()
print("switches %s | %d candidates: %.3f s wall, factor %.3f s, assembly %.3f s, %.3f ms per candidate | shared prefix %d, last-band cache %d | launches %d | checksum %.12e" % (
    [k for k in os.environ if k.startswith("GPCC_LARGE")], len(d4), dt, st["ms_factor"] * 1e-3, st["ms_assembly"] * 1e-3, dt * 1e3 / len(d4),
    st["n_shared_prefix"], st["n_tau_cache"], st["n_eval_launches"], float(np.sum(r["loglikel"]))))
