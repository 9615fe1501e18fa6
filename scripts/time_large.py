"""Dev script: tiled-path rates at a few (N, batch) points (logL only = blocked Cholesky, and logL+grad)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
ctx = gpcc_b200.Context(1, profiling=True)
for npb, batch in ((256, 512), (1024, 8), (1024, 64), (2048, 8), (2048, 64)):
    N = 3 * npb
    t, y, s, _ = gpcc_b200.synthetic_bands([npb] * 3, seed=5)
    var = np.array([a.var(ddof=1) for a in y])
    p = gpcc_b200.Problem(t, y, s, "matern52", ctx)
    rg = np.random.default_rng(0)
    d = np.zeros((batch, 3)); d[:, 1:] = rg.uniform(0, 8, (batch, 2))
    a = np.tile(np.sqrt(var) / 2.0, (batch, 1)); r = np.full(batch, 3.5)
    for grad in (False, True):
        p.loglik_batch(d, a, r, want_grad=grad)
        out = p.loglik_batch(d, a, r, want_grad=grad)
        st = ctx.stats()
        fl = batch * float(N) ** 3 * (1.0 if grad else 1 / 3)
        print("N=%5d batch %3d grad %d: %.3f ms/eval, %.2f TFLOP/s (%.0f %% of 37.0), info %d" % (
            N, batch, grad, st["ms_eval_kernels"] / batch, fl / (st["ms_factor"] * 1e-3) / 1e12, 100 * fl / (st["ms_factor"] * 1e-3) / 1e12 / 37.0, out[-1].max()), flush=True)
    p.close()
