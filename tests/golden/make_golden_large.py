"""Generates the expensive oracle fixtures (run from the repo root: python tests/golden/make_golden_large.py [what ...]).

  cfg3full   fit_cfg3_full.npz   BASELINE config 3 at FULL size: the oracle's L-BFGS optimum of all 101 x 101 = 10 201
                                 candidates (process pool; ~10 min on 8 cores) and the posterior
  cfg3nm     fit_cfg3_full_nm.npz  the same grid with the reference's optimiser (Nelder-Mead, g_tol 1e-6): ~1 h on 8 cores
  large      loglik_large.npz    fixed-hyper-parameter logL + gradient at N = 3072 (3 x 1024) and N = 6144 (3 x 2048,
                                 BASELINE config 4), two candidates each (one numpy potrf + potri at N = 6144 is ~1 min)

Like make_golden.py these pin the ORACLE (the reference cannot run here); make_golden.jl is the script a maintainer with
Julia + GPCC.jl runs to produce the same keys from the reference itself.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
_W = {}


def _init(tys, theta0, optimizer):
    from threadpoolctl import threadpool_limits
    _W["lim"] = threadpool_limits(1)
    _W["tys"], _W["theta0"], _W["opt"] = tys, theta0, optimizer


def _fit(dl):
    t, y, s = _W["tys"]
    r = oracle.gpcc(t, y, s, kernel="matern32", delays=dl, iterations=1000, rhomin=0.1, rhomax=300.0, theta0=_W["theta0"],
                    optimizer=_W["opt"], return_info=True)
    return r[0], r[3]["nfev"]


def cfg3_grid():
    t, y, s, d = oracle.simulatethreelightcurves()
    c = np.arange(0.0, 20.0001, 0.2)
    delays = np.array([[0.0, a, b] for b in c for a in c])            # d1 fastest (README.md:231-235)
    theta0, _ = oracle.initial_solutions(oracle.Problem(t, y, s, "matern32"), 1, 1, 5, 0.1, 300.0)
    return t, y, s, delays, theta0


def cfg3(optimizer, name, procs):
    t, y, s, delays, theta0 = cfg3_grid()
    t0 = time.time()
    with mp.get_context("fork").Pool(procs, initializer=_init, initargs=((t, y, s), theta0, optimizer)) as pool:
        out = pool.map(_fit, delays, chunksize=16)
    ll = np.array([o[0] for o in out]); nfev = np.array([o[1] for o in out])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), delays=delays, theta0=theta0[0], ll=ll, nfev=nfev,
                        post=oracle.getprobabilities(ll), optimizer=optimizer)
    print(name, "done in %.0f s, mean nfev %.1f, mode %s" % (time.time() - t0, nfev.mean(), delays[np.argmax(ll)]))


def large():
    out = {}
    for tag, nper, seed in (("n3072", [1024] * 3, 4), ("n6144", [2048] * 3, 4)):
        t, y, s, d = oracle.synthetic_bands(nper, seed=seed)
        p = oracle.Problem(t, y, s, "matern52")
        delays = np.array([[0.0, 2.0, 4.0], [0.0, 7.4, 12.2]])
        alpha, rho = np.tile([1.0, 2.2, 4.0], (2, 1)), np.array([3.5, 2.0])
        ll, grad = np.empty(2), np.empty((2, 4))
        for m in range(2):
            t0 = time.time()
            ll[m], grad[m] = p.loglik_grad(delays[m], alpha[m], rho[m])
            print(tag, m, ll[m], grad[m], "%.0f s" % (time.time() - t0), flush=True)
        out.update({tag + "_nper": np.array(nper), tag + "_seed": seed, tag + "_delays": delays, tag + "_alpha": alpha,
                    tag + "_rho": rho, tag + "_ll": ll, tag + "_grad": grad,
                    tag + "_tsum": np.array([a.sum() for a in t]), tag + "_ysum": np.array([a.sum() for a in y])})
    np.savez(os.path.join(OUT, "loglik_large.npz"), kernel="matern52", **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["large", "cfg3full"]
    procs = int(os.environ.get("GOLDEN_PROCS", "8"))
    if "large" in what:
        large()
    if "cfg3full" in what:
        cfg3("lbfgs", "fit_cfg3_full", procs)
    if "cfg3nm" in what:
        cfg3("neldermead", "fit_cfg3_full_nm", procs)
