"""Smallest case touching every kernel (fused path N=22 & N=150, tiled path N=260, postb, predict, posterior) for compute-sanitizer."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
from gpcc_b200 import Problem, Context
ctx = Context(1)
for nper, kern in (([7, 2, 13], "OU"), ([60, 50, 40], "matern32"), ([100, 90, 70], "matern52")):
    t, y, s, d = gpcc_b200.synthetic_bands(nper, seed=3, span=20.0)
    p = Problem(t, y, s, kern, ctx)
    L = len(nper)
    delays = np.zeros((3, L)); delays[:, 1:] = [[1.0] * (L - 1), [2.0] * (L - 1), [3.5] * (L - 1)]
    alpha = np.full((3, L), 1.3); rho = np.array([1.0, 3.0, 9.0])
    ll, g, info = p.loglik_batch(delays, alpha, rho, want_grad=True)
    ll2, info2 = p.loglik_batch(delays, alpha, rho)
    th = gpcc_b200.initial_solutions(y, 1, 1, 3, 0.1, 300.0)[0][0]
    r = p.grid_posterior(delays, th, iterations=5, rhomin=0.1, rhomax=300.0)
    mu, S = p.postb(delays[0], alpha[0], rho[0])
    tt = np.linspace(0, 20, 7)
    m_, sd_, Sf, _ = p.predict(delays[0], alpha[0], rho[0], [tt] * L, full_cov=True)
    tl, i3 = p.predict_loglik(delays[0], alpha[0], rho[0], [tt[:2]] * L, [np.full(2, 10.0)] * L, [np.full(2, 0.5)] * L)
    print(nper, kern, ll[:2], info, r["posterior"].sum(), mu[:2], sd_[:2], tl, i3)
