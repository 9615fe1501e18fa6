"""Generates tests/golden/pred_exact.npz: 50-digit mpmath values of postb and predictTest (oracle/exact.py follows
src/gpccfixdelay_marginaliseb.jl:248-250, :262-285, :303 literally) for
  case a  BASELINE config 1 (two light curves N = 110, matern32) at the oracle's fitted hyper-parameters
  case b  the 230-point synthetic problem of tests/test_gpu_parity.py::test_large_path_fit_and_postb_pred
Run from the repo root: python tests/golden/make_golden_exact.py   (~10 min: dense 50-digit inverses)
sigma_pred = sqrt((alpha^2 + Sigma_b) - k*'K^-1 k* + 1e-8) cancels ~4 digits, so float64 evaluations of the reference's formula
carry 1e-8..1e-6 relative noise; these values are what both the float64 oracle and the device library are held to.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402
from oracle.exact import postb_exact, predict_exact  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(tag, t, y, s, kernel, delays, alpha, rho, ttest, out):
    p = oracle.Problem(t, y, s, kernel)
    mu_b, S_b = postb_exact(p, delays, alpha, rho)
    mu, S, sd = predict_exact(p, delays, alpha, rho, ttest)
    out.update({tag + "_t": np.concatenate(t), tag + "_y": np.concatenate(y), tag + "_s": np.concatenate(s),
                tag + "_n": np.array([len(a) for a in t]), tag + "_kernel": kernel, tag + "_delays": np.asarray(delays, float),
                tag + "_alpha": np.asarray(alpha, float), tag + "_rho": float(rho),
                tag + "_ttest": np.concatenate(ttest), tag + "_ntest": np.array([len(a) for a in ttest]),
                tag + "_postb_mu": mu_b, tag + "_postb_Sigma": S_b, tag + "_pred_mu": mu, tag + "_pred_Sigma": S, tag + "_pred_sd": sd})
    om, oS = p.predict_full(delays, alpha, rho, ttest)
    osd = np.sqrt(np.maximum(np.diag(oS), 1e-6))
    print(tag, "float64 oracle vs exact: mu %.1e sd %.1e" % (np.max(np.abs(om - mu) / np.abs(mu)), np.max(np.abs(osd - sd) / sd)), flush=True)


def main():
    out = {}
    t3, y3, s3, d3 = oracle.simulatethreelightcurves()
    t2, y2, s2, d2 = t3[:2], y3[:2], s3[:2], d3[:2]
    theta0, _ = oracle.initial_solutions(oracle.Problem(t2, y2, s2, "matern32"), 1, 1, 5, 0.1, 300.0)
    r = oracle.gpcc(t2, y2, s2, kernel="matern32", delays=d2, iterations=1000, rhomax=300.0, theta0=theta0, optimizer="lbfgs")
    tt = np.arange(0.0, 20.0001, 0.5)
    case("a", t2, y2, s2, "matern32", d2, r[2][0], r[2][2], [tt[:21], tt[10:]], out)
    t, y, s, d = oracle.synthetic_bands([120, 110], seed=9)
    theta0, _ = oracle.initial_solutions(oracle.Problem(t, y, s, "matern32"), 1, 1, 5, 0.1, 300.0)
    r = oracle.gpcc(t, y, s, kernel="matern32", delays=[0.0, 2.0], iterations=1000, rhomax=300.0, theta0=theta0, optimizer="lbfgs")
    tt = np.linspace(0.0, 40.0, 33)
    case("b", t, y, s, "matern32", [0.0, 2.0], r[2][0], r[2][2], [tt, tt], out)
    np.savez_compressed(os.path.join(OUT, "pred_exact.npz"), **out)


if __name__ == "__main__":
    main()
