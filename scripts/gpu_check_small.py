"""First GPU check of the fused small-N kernel against the oracle (dev script, not a test)."""
import time, sys
import numpy as np
sys.path.insert(0, ".")
import oracle, gpcc_b200
from gpcc_b200 import Problem, Context

ctx = Context(1, profiling=True)
for kern in ["matern32", "OU", "rbf", "matern52"]:
    for nb in (2, 3):
        t, y, s, d = oracle.simulatethreelightcurves()
        t, y, s, d = t[:nb], y[:nb], s[:nb], d[:nb]
        op = oracle.Problem(t, y, s, kern)
        p = Problem(t, y, s, kern, ctx)
        rg = np.random.default_rng(0)
        M = 64
        delays = np.zeros((M, nb)); delays[:, 1:] = rg.uniform(0, 10, (M, nb - 1))
        alpha = rg.uniform(0.5, 3.0, (M, nb)); rho = rg.uniform(0.5, 20, M)
        ll, g, info = p.loglik_batch(delays, alpha, rho, want_grad=True)
        ref = [op.loglik_grad(delays[m], alpha[m], rho[m]) for m in range(M)]
        rl = np.array([r[0] for r in ref]); rgd = np.array([r[1] for r in ref])
        print(kern, nb, "N", op.N, "ll relerr %.2e" % np.max(np.abs(ll - rl) / np.abs(rl)),
              "grad relerr %.2e" % (np.max(np.abs(g - rgd)) / np.max(np.abs(rgd))), "info", info.max(), flush=True)

# non-PD detection: zero noise + duplicated points
t, y, s, d = oracle.simulatetwolightcurves()
p = Problem(t, y, s, "matern32", ctx)
ll, info = p.loglik_batch([[0, 2.0]], [[1.0, 1.0]], [1e-9])
print("tiny rho", ll, info)

# throughput of the raw kernel
t, y, s, d = oracle.simulatethreelightcurves()
p = Problem(t, y, s, "matern32", ctx)
M = 148 * 64
rg = np.random.default_rng(1)
delays = np.zeros((M, 3)); delays[:, 1:] = rg.uniform(0, 20, (M, 2))
alpha = rg.uniform(0.5, 3.0, (M, 3)); rho = rg.uniform(0.5, 20, M)
for it in range(3):
    t0 = time.time(); p.loglik_batch(delays, alpha, rho, want_grad=True); dt = time.time() - t0
    st = ctx.stats()
    print("N=150 batch", M, "wall %.2f ms kernel %.2f ms -> %.1f us/eval/SM-slot, %.2f TFLOP/s (N^3)" % (
        dt * 1e3, st["ms_eval_kernels"], st["ms_eval_kernels"] * 1e3 / (M / 148), M * 150.0**3 / st["ms_eval_kernels"] / 1e9))
t, y, s, d = oracle.simulatetwolightcurves()
p2 = Problem(t, y, s, "matern32", ctx)
for it in range(3):
    t0 = time.time(); p2.loglik_batch(delays[:, :2], alpha[:, :2], rho, want_grad=True); dt = time.time() - t0
    st = ctx.stats()
    print("N=110 batch", M, "wall %.2f ms kernel %.2f ms, %.2f TFLOP/s (N^3)" % (dt * 1e3, st["ms_eval_kernels"], M * 110.0**3 / st["ms_eval_kernels"] / 1e9))

# fit on cfg2-like grid
t, y, s, d = oracle.simulatetwolightcurves()
cands = np.arange(0, 10.0001, 0.1)
delays = np.stack([np.zeros_like(cands), cands], 1)
theta0 = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
for it in range(2):
    t0 = time.time()
    res = p2.grid_posterior(delays, theta0, iterations=1000, rhomin=0.1, rhomax=300.0)
    dt = time.time() - t0
    print("cfg2 grid: %.1f ms, nfev mean %.1f max %d, status" % (dt * 1e3, res["nfev"].mean(), res["nfev"].max()), np.bincount(res["info"] + 1), ctx.stats())
print("posterior mode at", cands[np.argmax(res["posterior"])], "sum", res["posterior"].sum())
# compare a few candidates with the oracle L-BFGS / NM
for m in [0, 20, 50, 99]:
    r = oracle.gpcc(t, y, s, kernel="matern32", delays=delays[m], iterations=1000, rhomax=300.0, theta0=theta0[None], optimizer="lbfgs")
    r2 = oracle.gpcc(t, y, s, kernel="matern32", delays=delays[m], iterations=1000, rhomax=300.0, theta0=theta0[None])
    print(m, "gpu %.8f oracle-lbfgs %.8f oracle-nm %.8f" % (res["loglikel"][m], r[0], r2[0]))
