"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/gpcc_b200.h declares,
and refuses to work without a CUDA device (no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import gpcc_b200
from gpcc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "gpcc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpcc_[a-z_0-9]+)\s*\(", txt)))


def test_library_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = _lib.load()
    assert lib.gpcc_version() >= 100


def test_every_header_symbol_is_exported():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 19
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == syms, "python binding and header disagree"


def test_library_does_not_link_oracle_or_cpu_blas():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "blas" not in out.lower() and "lapack" not in out.lower()
    # the product package never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "gpcc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_a_gpu():
    with pytest.raises(gpcc_b200.GpccError) as e:
        gpcc_b200.Context(1)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.gpcc_ctx_create(1, None, C.byref(h)) != 0 and not h.value
    assert lib.gpcc_last_error()


def test_fit_options_default_matches_reference_constants():
    lib = _lib.load()
    o = _lib.FitOptions()
    assert lib.gpcc_fit_options_default(C.byref(o)) == 0
    assert o.max_iter == 1000 and o.rhomin == 0.1 and o.alpha_floor == 1e-8       # :46, :112
    assert o.history == 8 and o.transform_id == 0 and o.theta0_per_candidate == 0
    assert o.optimizer == 0 and o.nm_gtol == 1e-6                                   # L-BFGS by default; g_tol of :205 for the NM option
    assert lib.gpcc_fit_options_default(None) != 0


def test_null_and_invalid_arguments_are_rejected_without_a_device():
    lib = _lib.load()
    assert lib.gpcc_ctx_destroy(None) == 0 and lib.gpcc_problem_destroy(None) == 0
    assert lib.gpcc_ctx_device_count(None) == 0
    out = C.c_void_p()
    assert lib.gpcc_problem_create(None, 2, None, None, None, None, 0, None, None, C.byref(out)) < 0
    assert lib.gpcc_loglik_batch(None, 1, None, None, None, 0, None, None, None) < 0
    assert lib.gpcc_fit_batch(None, 1, None, 1, None, None, None, None, None, None, None, None, None) < 0
    assert lib.gpcc_postb(None, None, None, 1.0, None, None) < 0
    assert b"NULL" in lib.gpcc_last_error()


def test_host_side_mirror_of_reference_api():
    # kernels are identity objects mapped to the enum; anything else is an error (no CPU path)
    assert [k.kid for k in (gpcc_b200.OU, gpcc_b200.rbf, gpcc_b200.matern32, gpcc_b200.matern52)] == [0, 1, 2, 3]
    from gpcc_b200.api import _kernel_id
    assert _kernel_id("matern32") == 2
    with pytest.raises(gpcc_b200.GpccError):
        _kernel_id(lambda a, b: 1.0)
    pr = gpcc_b200.uniformpriordelay(L=1e44, z=0.5)                        # uniformpriordelay.jl:12
    assert pr.a == 0.0 and pr.b == pytest.approx(10 ** 1.559 * 1.5)
    lp = pr.logpdf([-1.0, 1.0, 1e3])
    assert lp[0] == -np.inf and lp[2] == -np.inf and lp[1] == pytest.approx(-np.log(pr.b))


def test_initial_solutions_follow_reference_draw_order():
    import oracle
    t, y, s, d = oracle.simulatetwolightcurves()
    th, rho0 = gpcc_b200.initial_solutions(y, seed=1, numberofrestarts=1, initialrandom=5, rhomin=0.1, rhomax=300.0)
    tho, rho0o = oracle.initial_solutions(oracle.Problem(t, y, s, "OU"), 1, 1, 5, 0.1, 300.0)
    assert th.shape == (1, 5, 3) and np.array_equal(th, tho) and np.array_equal(rho0, rho0o)
    assert 0.101 <= rho0[0] <= 299.999                                                       # :166
    var = np.array([a.var(ddof=1) for a in y])
    a0 = oracle.makepositive(th[0, :, :2])
    assert np.all(a0 >= 0.8 * var - 1e-9) and np.all(a0 <= 1.2 * var + 1e-9)                 # :188
    assert np.all(th[0, :, 2] == th[0, 0, 2])                                                # rho0 shared by the draws (:196)
    th3, r3 = gpcc_b200.initial_solutions(y, seed=1, numberofrestarts=4, initialrandom=2, rhomin=0.1, rhomax=20.0)
    assert np.allclose(np.diff(np.log(r3)), np.log(r3[1] / r3[0]))                           # log-spaced grid (:172)


def test_repaired_logpdf_matches_the_reference_recipe():
    """Host-side repair of a predictive covariance that is not positive definite (reference :323-341): eigenvalues clamped
    at 1e-6, then the Gaussian log-density -- checked against the oracle's restatement of the same branch."""
    import numpy as np
    import gpcc_b200
    rg = np.random.default_rng(3)
    Q, _ = np.linalg.qr(rg.normal(size=(6, 6)))
    S = (Q * np.array([2.0, 1.0, 0.5, 1e-3, -1e-9, -0.2])) @ Q.T          # indefinite
    mu, y = rg.normal(size=6), rg.normal(size=6)
    got = gpcc_b200.api.repaired_logpdf(mu, S, y)
    w, V = np.linalg.eigh(0.5 * (S + S.T))
    Sr = (V * np.maximum(w, 1e-6)) @ V.T
    sign, logdet = np.linalg.slogdet(Sr)
    ref = -0.5 * (6 * np.log(2 * np.pi) + logdet + (y - mu) @ np.linalg.solve(Sr, y - mu))
    assert sign > 0 and abs(got - ref) < 1e-9 * abs(ref)
