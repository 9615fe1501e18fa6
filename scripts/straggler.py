import sys, numpy as np
sys.path.insert(0, ".")
import gpcc_b200, oracle
ctx = gpcc_b200.Context(1)
t, y, s, d = gpcc_b200.simulatethreelightcurves()
p = gpcc_b200.Problem(t, y, s, "matern32", ctx)
theta0 = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
W = 2
c2 = np.arange(0.0, 20.0001, 0.2); c3 = np.linspace(0.0, 20.0, 101 * W)
delays_all = np.array([[0.0, a, b] for b in c3 for a in c2])
delays = np.ascontiguousarray(delays_all[1::W])
r = p.fit_batch(delays, theta0, iterations=1000, rhomin=0.1, rhomax=300.0)
k = int(np.argmax(r["nfev"]))
print("straggler", k, delays[k], "nfev", r["nfev"][k], "iters", r["iters"][k], "ll", r["loglikel"][k], "alpha", r["alpha"][k], "rho", r["rho"][k], "theta", r["theta"][k], "best ll in shard", r["loglikel"].max())
for it in (10, 30, 100, 300, 1000, 3000):
    q = p.fit_batch(delays[k:k+1], theta0, iterations=it, rhomin=0.1, rhomax=300.0)
    ll, g, info = p.loglik_theta_batch(delays[k:k+1], q["theta"], 0.1, 300.0, want_grad=True)
    print(it, "ll %.10f" % q["loglikel"][0], "theta", q["theta"][0], "grad_theta", g[0], "status", q["info"][0], "nfev", q["nfev"][0])
o = oracle.gpcc(t, y, s, kernel="matern32", delays=delays[k], iterations=1000, rhomax=300.0, theta0=theta0[None], optimizer="lbfgs", return_info=True)
print("oracle lbfgs", o[0], o[3])
o = oracle.gpcc(t, y, s, kernel="matern32", delays=delays[k], iterations=1000, rhomax=300.0, theta0=theta0[None], return_info=True)
print("oracle NM", o[0], o[3])
