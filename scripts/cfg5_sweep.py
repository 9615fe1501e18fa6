"""BASELINE config 5: kernel sweep OU/rbf/matern32/matern52 x N in {64,256,1024,4096} per band (L=3), batched
fixed-delay evaluations at fixed theta -- the batch-size / N crossover study.  Writes gpurun_out/cfg5_sweep_r2.json."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
from gpcc_b200 import Problem, Context
ctx = Context(1, profiling=True)
rows = []
for npb in (64, 256, 1024, 4096):
    N = 3 * npb
    t, y, s, d = gpcc_b200.synthetic_bands([npb] * 3, seed=5)
    var = np.array([a.var(ddof=1) for a in y])
    for kern in ("OU", "rbf", "matern32", "matern52"):
        if npb >= 1024 and kern not in ("matern52",):
            continue
        p = Problem(t, y, s, kern, ctx)
        for batch in (1, 8, 64, 512, 4096):
            if N >= 12288 and batch > 16: continue
            if N >= 3072 and batch > 64: continue
            if N >= 768 and batch > 512: continue
            rg = np.random.default_rng(0)
            delays = np.zeros((batch, 3)); delays[:, 1:] = rg.uniform(0, 8, (batch, 2))
            alpha = np.tile(np.sqrt(var) / 2.0, (batch, 1)); rho = np.full(batch, 3.5)
            for grad in (False, True):
                p.loglik_batch(delays, alpha, rho, want_grad=grad)
                t0 = time.perf_counter(); out = p.loglik_batch(delays, alpha, rho, want_grad=grad); dt = time.perf_counter() - t0
                st = ctx.stats()
                flop = batch * float(N) ** 3 * (1.0 if grad else 1.0 / 3.0)      # logL only = forward elimination / blocked Cholesky on both paths
                rows.append(dict(kernel=kern, n_per_band=npb, N=N, batch=batch, grad=grad, path=["fused-register", "tiled-DMMA"][st["path"]],
                                 wall_ms=dt * 1e3, kernel_ms=st["ms_eval_kernels"], evals_per_s=batch / dt,
                                 tflops=flop / (st["ms_eval_kernels"] * 1e-3) / 1e12, info_max=int(np.max(out[-1]))))
                print(rows[-1], flush=True)
        p.close()
json.dump(dict(what="cfg5 sweep on one B200, fixed theta (alpha = sd(y)/2, rho = 3.5); flop model: N^3 for logL+grad (symmetric sweep), N^3/3 for logL only (forward elimination / blocked Cholesky), both paths",
               rows=rows), open("gpurun_out/cfg5_sweep_r2.json", "w"), indent=1)
