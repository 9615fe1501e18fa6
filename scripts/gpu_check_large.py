"""GPU check of the tiled large-N path against the oracle + first timings (dev script)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle, gpcc_b200
from gpcc_b200 import Problem, Context
ctx = Context(1, profiling=True)

def check(nper, kernel, M=3, span=None, seed=3):
    t, y, s, d = gpcc_b200.synthetic_bands(nper, seed=seed, span=span)
    op = oracle.Problem(t, y, s, kernel); p = Problem(t, y, s, kernel, ctx)
    L = len(nper)
    rg = np.random.default_rng(1)
    delays = np.zeros((M, L)); delays[:, 1:] = rg.uniform(0, 6, (M, L - 1))
    alpha = rg.uniform(0.5, 2.5, (M, L)); rho = rg.uniform(1.0, 8.0, M)
    t0 = time.time(); ll, g, info = p.loglik_batch(delays, alpha, rho, want_grad=True); t1 = time.time() - t0
    st = ctx.stats()
    ll0, info0 = p.loglik_batch(delays, alpha, rho)
    st0 = ctx.stats()
    ref = [op.loglik_grad(delays[m], alpha[m], rho[m]) for m in range(M)]
    rl = np.array([r[0] for r in ref]); rgd = np.array([r[1] for r in ref])
    print(nper, kernel, "N", op.N, "path", st["path"], "info", info, info0,
          "sweep relerr %.2e fwd relerr %.2e grad relerr %.2e" % (np.max(np.abs(ll - rl) / np.abs(rl)), np.max(np.abs(ll0 - rl) / np.abs(rl)),
                                                               np.max(np.abs(g - rgd) / np.max(np.abs(rgd), axis=1, keepdims=True))),
          "ms sweep %.1f (asm %.2f fac %.1f grad %.2f) fwd %.1f launches %d" % (st["ms_eval_kernels"], st["ms_assembly"], st["ms_factor"], st["ms_gradreduce"], st0["ms_eval_kernels"], st["n_eval_launches"]), flush=True)
    return p, op, delays, alpha, rho

check([100, 90, 70], "matern32")          # N=260, padded to 384
check([256, 256, 256], "matern52")        # N=768
check([300, 212], "OU")                   # N=512 exactly 4 tiles
p, op, delays, alpha, rho = check([256, 256, 256], "rbf", M=2)
# postb / predict through the large path
mu, S = p.postb(delays[0], alpha[0], rho[0]); omu, oS = op.postb(delays[0], alpha[0], rho[0])
print("postb relerr", np.max(np.abs(mu - omu) / np.abs(omu)), np.max(np.abs(S - oS)) / np.max(np.abs(oS)))
tt = np.linspace(0, 100, 50)
m_, sd_, _, _ = p.predict(delays[0], alpha[0], rho[0], [tt] * 3); om, osd = op.predict(delays[0], alpha[0], rho[0], tt)
print("pred relerr", np.max(np.abs(m_ - np.concatenate(om)) / np.abs(np.concatenate(om))), np.max(np.abs(sd_ - np.concatenate(osd)) / np.concatenate(osd)))

# cfg4 size: 3 x 2048, matern52
for nper, M in (([1024, 1024, 1024], 8), ([2048, 2048, 2048], 8)):
    t, y, s, d = gpcc_b200.synthetic_bands(nper, seed=4)
    N = sum(nper)
    p = Problem(t, y, s, "matern52", ctx)
    rg = np.random.default_rng(2)
    delays = np.zeros((M, 3)); delays[:, 1:] = rg.uniform(0, 19.8, (M, 2))
    alpha = np.tile([1.0, 2.2, 4.0], (M, 1)); rho = np.full(M, 3.5)
    for mode in (False, True):
        for it in range(2):
            t0 = time.time(); out = p.loglik_batch(delays, alpha, rho, want_grad=mode); dt = time.time() - t0
        st = ctx.stats()
        flops = M * float(N) ** 3 * (1.0 if mode else 1.0 / 3.0)
        print("N=%d M=%d grad=%d: wall %.1f ms kernels %.1f ms (asm %.2f fac %.1f grad %.2f) -> %.2f TFLOP/s algorithmic; asm %.0f GB/s; ll[0]=%.6f info %s" % (
            N, M, mode, dt * 1e3, st["ms_eval_kernels"], st["ms_assembly"], st["ms_factor"], st["ms_gradreduce"], flops / st["ms_factor"] / 1e9,
            M * 4.0 * N * (N + 1) / st["ms_assembly"] / 1e6, out[0][0], out[-1][:3]), flush=True)
    if N <= 3072:
        op = oracle.Problem(t, y, s, "matern52")
        rl, rgd = op.loglik_grad(delays[0], alpha[0], rho[0])
        print("  oracle ll %.6f relerr %.2e grad relerr %.2e" % (rl, abs(out[0][0] - rl) / abs(rl), np.max(np.abs(out[1][0] - rgd)) / np.max(np.abs(rgd))))
