"""Dev script: one fitted candidate on the tiled path where the device optimum was 1.3e-6 below the oracle's (random-shape test)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200, oracle
rg = np.random.default_rng(2026)
L = int(rg.integers(1, 6)); nper = [int(v) for v in rg.integers(2, 91, L)]
nper = [int(v) for v in rg.integers(60, 120, 3)]; L = 3
t, y, s, d = gpcc_b200.synthetic_bands(nper, seed=100, span=float(rg.uniform(8.0, 40.0)))
M = 3
delays = np.zeros((M, L)); delays[:, 1:] = rg.uniform(-4, 9, (M, L - 1))
p = gpcc_b200.Problem(t, y, s, "OU")
th = gpcc_b200.initial_solutions(y, 1, 1, 3, 0.1, 80.0)[0][0]
for kw in (dict(), dict(gtol=1e-9), dict(ftol=1e-15), dict(gtol=1e-9, ftol=1e-15)):
    r = p.fit_batch(delays[:1], th, iterations=400, rhomin=0.1, rhomax=80.0, **kw)
    ll, g, info = p.loglik_theta_batch(delays[:1], r["theta"], 0.1, 80.0, want_grad=True)
    print(kw, "ll %.10f iters %d nfev %d status %d |g|inf %.2e theta %s" % (r["loglikel"][0], r["iters"][0], r["nfev"][0], r["info"][0], np.max(np.abs(g)), r["theta"][0]))
o = oracle.gpcc(t, y, s, kernel="OU", delays=delays[0], iterations=400, rhomin=0.1, rhomax=80.0, theta0=th[None], optimizer="lbfgs", return_info=True)
print("oracle ll %.10f nfev %d theta %s" % (o[0], o[3]["nfev"], o[3]["theta"]))
op = oracle.Problem(t, y, s, "OU")
print("oracle grad at device theta", op.objective_grad_theta(r["theta"][0], delays[0], 0.1, 80.0))
