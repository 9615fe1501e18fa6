"""Short workload for ncu: the tiled path at a mid size (3 x 256 points, N = 768, 512 matrices, logL only = blocked Cholesky)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
npb, batch = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 512)
ctx = gpcc_b200.Context(1, profiling=True)
t, y, s, _ = gpcc_b200.synthetic_bands([npb] * 3, seed=5)
var = np.array([a.var(ddof=1) for a in y])
p = gpcc_b200.Problem(t, y, s, "matern52", ctx)
rg = np.random.default_rng(0)
d = np.zeros((batch, 3)); d[:, 1:] = rg.uniform(0, 8, (batch, 2))
a = np.tile(np.sqrt(var) / 2.0, (batch, 1)); r = np.full(batch, 3.5)
for rep in range(2):
    out = p.loglik_batch(d, a, r)
    st = ctx.stats()
    fl = batch * float(3 * npb) ** 3 / 3
    print("N=%d batch %d: factor %.3f ms, %.2f TFLOP/s, launches %d" % (3 * npb, batch, st["ms_factor"], fl / (st["ms_factor"] * 1e-3) / 1e12, st["n_eval_launches"]), flush=True)
