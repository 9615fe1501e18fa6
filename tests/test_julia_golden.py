"""Consumes tests/golden/julia_*.npz -- the outputs of the REAL GPCC.jl written by tests/golden/make_golden.jl where Julia,
GPCC.jl and MiscUtil exist.  Julia is not available in the build image, so these files are normally absent and every test
here skips; once a maintainer has run the script they pin the oracle (CPU) and, under -m gpu, the CUDA library against
the reference itself: MiscUtil's transforms, the fixed-hyper-parameter log-likelihood (1e-10 relative), and the cfg1 fit /
cfg2 grid started from the reference's own MersenneTwister draws with the reference's own optimiser (Nelder-Mead)."""
import glob
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, load_golden


def _julia(name):
    path = os.path.join(GOLDEN, "julia_" + name + ".npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/julia_%s.npz absent: run tests/golden/make_golden.jl with Julia + GPCC.jl" % name)
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def test_miscutil_transforms_are_the_presumed_ones():
    j = _julia("miscutil")
    assert np.allclose(oracle.makepositive(j["x"]), j["makepositive"], rtol=1e-14, atol=1e-300)
    assert np.allclose(oracle.invmakepositive(j["makepositive"]), j["invmakepositive"], rtol=1e-10, atol=1e-12)
    lo, hi = float(j["lo"]), float(j["hi"])
    assert np.allclose(oracle.transformbetween(j["x"], lo, hi), j["transformbetween"], rtol=1e-14)
    assert np.allclose(oracle.invtransformbetween(j["transformbetween"], lo, hi), j["invtransformbetween"], rtol=1e-9, atol=1e-9)


def _loglik_cases():
    return sorted(os.path.basename(p)[len("julia_"):-4] for p in glob.glob(os.path.join(GOLDEN, "julia_loglik_*.npz")))


def test_oracle_loglik_matches_reference():
    cases = _loglik_cases()
    if not cases:
        pytest.skip("no tests/golden/julia_loglik_*.npz: run tests/golden/make_golden.jl")
    for name in cases:
        g, j = load_golden(name), _julia(name)
        assert np.max(np.abs(g["loglik"] - j["loglik"]) / np.abs(j["loglik"])) < 1e-10, name
        assert np.allclose(g["mub"], j["mub"], rtol=1e-13) and np.allclose(g["Sigmab"], j["Sigmab"], rtol=1e-12)


def test_oracle_fit_and_predictions_match_reference():
    j, g = _julia("fit_cfg1"), load_golden("fit_cfg1_cfg2")
    r = oracle.gpcc(g["tb"], g["yb"], g["sb"], kernel="matern32", delays=g["truedelays"], iterations=1000, rhomax=300.0,
                    theta0=j["theta0"][None], optimizer="neldermead")
    assert abs(r[0] - float(j["loglikel"])) < 1e-6
    p = oracle.Problem(g["tb"], g["yb"], g["sb"], "matern32")
    mu, S = p.postb(g["truedelays"], j["alpha"], float(j["rho"]))
    assert np.allclose(mu, j["postb_mu"], rtol=1e-8) and np.allclose(S, j["postb_Sigma"], rtol=1e-8)
    om, osd = p.predict(g["truedelays"], j["alpha"], float(j["rho"]), g["ttest"])
    assert np.allclose(np.array(om), j["pred_mu"], rtol=1e-8) and np.allclose(np.array(osd), j["pred_sd"], rtol=1e-7)


@pytest.mark.gpu
def test_library_matches_reference_outputs():
    import gpcc_b200
    _julia("fit_cfg1")                       # skip before touching the GPU when the reference outputs are absent
    ctx = gpcc_b200.Context(1)
    for name in _loglik_cases():
        g, j = load_golden(name), _julia(name)
        p = gpcc_b200.Problem(g["tb"], g["yb"], g["sb"], g["kernel"], ctx)
        ll, info = p.loglik_batch(g["delays"], g["alpha"], g["rho"])
        assert np.max(np.abs(ll - j["loglik"]) / np.abs(j["loglik"])) < 1e-10, name
    j, g = _julia("fit_cfg1"), load_golden("fit_cfg1_cfg2")
    for optimizer, tol in (("neldermead", 1e-6), ("lbfgs", 2e-5)):        # NM stops up to ~1e-5 short of the optimum L-BFGS reaches
        ll, pred, (alpha, postb, rho) = gpcc_b200.gpcc(g["tb"], g["yb"], g["sb"], kernel=gpcc_b200.matern32, delays=g["truedelays"],
                                                      iterations=1000, rhomax=300, theta0=j["theta0"], ctx=ctx, verbose=False, optimizer=optimizer)
        assert abs(ll - float(j["loglikel"])) < tol, optimizer
    p = gpcc_b200.Problem(g["tb"], g["yb"], g["sb"], "matern32", ctx)
    st = p.fit_state(g["truedelays"], j["alpha"], float(j["rho"]))
    mu, S = st.postb()
    assert np.allclose(mu, j["postb_mu"], rtol=1e-8) and np.allclose(S, j["postb_Sigma"], rtol=1e-8)
    m_, sd_, _, _ = st.predict([g["ttest"]] * 2)
    assert np.allclose(m_.reshape(2, -1), j["pred_mu"], rtol=1e-8) and np.allclose(sd_.reshape(2, -1), j["pred_sd"], rtol=1e-7)
    jg = _julia("grid_cfg2")
    delays = np.stack([np.zeros_like(jg["cands"]), jg["cands"]], 1)
    r = p.grid_posterior(delays, j["theta0"], iterations=1000, rhomin=0.1, rhomax=300.0, optimizer="neldermead")
    assert np.max(np.abs(r["posterior"] - jg["post_flat"])) < 1e-4
    assert np.max(np.abs(r["loglikel"] - jg["ll_grid"])[jg["post_flat"] > 1e-12]) < 1e-6
