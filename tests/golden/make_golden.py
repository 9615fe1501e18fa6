"""Generates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference itself cannot run here (no Julia, un-vendored dependencies) and ships no golden vectors, so
these fixtures pin the ORACLE (numpy/scipy restatement of the reference, see oracle/__init__.py), which is
in turn validated against mpmath / finite differences / Woodbury in tests/test_oracle.py.
Sizes follow BASELINE.json configs 1-3 and 5 (small-N part).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pack_bands(arrs):
    return np.concatenate(arrs), np.array([len(a) for a in arrs])


def fixed_theta_case(name, t, y, s, kernel, M, seed, delay_hi=10.0):
    p = oracle.Problem(t, y, s, kernel)
    rg = np.random.default_rng(seed)
    L = p.L
    delays = np.zeros((M, L)); delays[:, 1:] = rg.uniform(0.0, delay_hi, (M, L - 1))
    alpha = rg.uniform(0.3, 3.0, (M, L)); rho = np.exp(rg.uniform(np.log(0.2), np.log(100.0), M))
    ll = np.empty(M); grad = np.empty((M, L + 1))
    for m in range(M):
        ll[m], grad[m] = p.loglik_grad(delays[m], alpha[m], rho[m])
    tt, n = pack_bands(t); yy, _ = pack_bands(y); ss, _ = pack_bands(s)
    np.savez(os.path.join(OUT, name + ".npz"), t=tt, y=yy, s=ss, n=n, kernel=kernel, delays=delays, alpha=alpha, rho=rho,
             loglik=ll, grad=grad, mub=p.mub, Sigmab=p.Sigmab)
    print(name, "N", p.N, "M", M, "ll range", ll.min(), ll.max())


def main():
    t3, y3, s3, d3 = oracle.simulatethreelightcurves()
    t2, y2, s2, d2 = t3[:2], y3[:2], s3[:2], d3[:2]
    # fixed-hyper-parameter log-likelihood + gradient (tolerance 1e-10 relative), cfg1/2 (2 bands) and cfg3 (3 bands)
    for k in oracle.KERNELS:
        fixed_theta_case(f"loglik_2band_{k}", t2, y2, s2, k, 24, 11)
        fixed_theta_case(f"loglik_3band_{k}", t3, y3, s3, k, 24, 12, delay_hi=20.0)
    # cfg5 small-N member: 3 bands x 64 points (N=192, the largest size of the fused kernel)
    t, y, s, d = oracle.synthetic_bands([64, 64, 64], seed=5)
    fixed_theta_case("loglik_3x64_matern52", t, y, s, "matern52", 8, 13, delay_hi=8.0)
    # ragged / tiny bands
    t, y, s, d = oracle.synthetic_bands([7, 2, 13], seed=6, span=10.0)
    fixed_theta_case("loglik_ragged_OU", t, y, s, "OU", 8, 14, delay_hi=5.0)
    t, y, s, d = oracle.synthetic_bands([9], seed=7, span=10.0)
    fixed_theta_case("loglik_single_band_rbf", t, y, s, "rbf", 4, 15)

    # cfg1: single fit at the true delays; cfg2: 1-D grid 0:0.1:10 (every 5th candidate stored to keep it small)
    theta0, rho0 = oracle.initial_solutions(oracle.Problem(t2, y2, s2, "matern32"), 1, 1, 5, 0.1, 300.0)
    r = oracle.gpcc(t2, y2, s2, kernel="matern32", delays=d2, iterations=1000, rhomax=300.0, theta0=theta0,
                    optimizer="lbfgs", return_info=True)
    rnm = oracle.gpcc(t2, y2, s2, kernel="matern32", delays=d2, iterations=1000, rhomax=300.0, theta0=theta0)
    ttest = np.arange(0.0, 20.0001, 0.5)
    mu, sd = r[1](ttest)
    mufull, Sfull = r[1]([ttest[:7], ttest[5:9]])
    tl = r[1]([[9.0, 10.0, 11.0], [9.0, 10.0, 11.0]], [[6.34, 5.49, 5.38], [13.08, 12.37, 15.69]],
              [[0.34, 0.42, 0.2], [0.87, 0.8, 0.66]])
    cands = np.arange(0.0, 10.0001, 0.1)
    ll_grid = np.array([oracle.gpcc(t2, y2, s2, kernel="matern32", delays=[0.0, c], iterations=1000, rhomax=300.0,
                                    theta0=theta0, optimizer="lbfgs")[0] for c in cands])
    ll_grid_nm = np.array([oracle.gpcc(t2, y2, s2, kernel="matern32", delays=[0.0, c], iterations=1000, rhomax=300.0,
                                       theta0=theta0)[0] for c in cands])
    prior = oracle.uniformpriordelay(L=1e44, z=0.0)
    tt, n = pack_bands(t2); yy, _ = pack_bands(y2); ss, _ = pack_bands(s2)
    np.savez(os.path.join(OUT, "fit_cfg1_cfg2.npz"), t=tt, y=yy, s=ss, n=n, theta0=theta0[0], truedelays=d2,
             loglikel=r[0], loglikel_nm=rnm[0], alpha=r[2][0], rho=r[2][2], postb_mu=r[2][1][0], postb_Sigma=r[2][1][1],
             ttest=ttest, pred_mu=np.array(mu), pred_sd=np.array(sd), predfull_mu=mufull, predfull_Sigma=Sfull,
             test_loglik=tl, cands=cands, ll_grid=ll_grid, ll_grid_nm=ll_grid_nm,
             post_flat=oracle.getprobabilities(ll_grid), post_prior=oracle.getprobabilities(ll_grid, prior.logpdf(cands)),
             prior_upper=prior.b)
    print("cfg1 loglikel", r[0], "nm", rnm[0], "grid mode", cands[np.argmax(ll_grid)],
          "max |lbfgs-nm|", np.max(np.abs(ll_grid - ll_grid_nm)))

    # cfg3: coarse 2-D sub-grid of (0:0.2:20)^2 (every 10th point: 11 x 11 = 121 candidates), d1 fastest (README.md:231-235)
    theta0_3, _ = oracle.initial_solutions(oracle.Problem(t3, y3, s3, "matern32"), 1, 1, 5, 0.1, 300.0)
    c1 = np.arange(0.0, 20.0001, 2.0)
    delays3 = np.array([[0.0, a, b] for b in c1 for a in c1])
    ll3 = np.array([oracle.gpcc(t3, y3, s3, kernel="matern32", delays=dl, iterations=1000, rhomax=300.0, theta0=theta0_3,
                                optimizer="lbfgs")[0] for dl in delays3])
    tt, n = pack_bands(t3); yy, _ = pack_bands(y3); ss, _ = pack_bands(s3)
    np.savez(os.path.join(OUT, "fit_cfg3_subgrid.npz"), t=tt, y=yy, s=ss, n=n, theta0=theta0_3[0], delays=delays3, ll=ll3,
             post=oracle.getprobabilities(ll3))
    print("cfg3 subgrid mode", delays3[np.argmax(ll3)])


if __name__ == "__main__":
    main()
