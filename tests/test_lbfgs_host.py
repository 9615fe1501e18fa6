"""The L-BFGS state machine of gpcc_b200/csrc/lbfgs.h on the HOST (it is plain C++ behind GPCC_HD; the persistent fit kernel
and the host-driven batched loop run this same code): compiled with g++ and driven on two analytic problems."""
import os, shutil, subprocess, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lines(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("lbfgs") / "lbfgs_host_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "gpcc_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "lbfgs_host_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    return {l.split()[0]: [float(v) for v in l.split()[1:]] for l in out}


def test_rosenbrock_converges(lines):
    nfev, iters, status, f, *x = lines["rosenbrock"]
    assert status == 0 and f < 1e-12 and np.allclose(x, 1.0, atol=1e-5) and nfev < 200


def test_relative_step_cap_of_the_scales(lines):
    """A scale whose optimum lies far up the linear branch of the softplus (alpha = 150: the edge-of-grid candidates of the
    three-band grid): with the fixed cap of 2 units per iteration the optimiser walks there; with the cap growing as
    0.25 theta_i above the knee (LbfgsOptions::n_scale) it gets there geometrically -- same optimum, far fewer evaluations."""
    n0, i0, s0, f0, *x0 = lines["scales_fixed_cap"]
    n1, i1, s1, f1, *x1 = lines["scales_relative_cap"]
    assert s0 == 0 and s1 == 0 and f0 < 1e-12 and f1 < 1e-12
    assert np.allclose(x0, x1, rtol=1e-5, atol=1e-5) and abs(x1[2] - 150.0) < 1e-3
    assert n0 > 70 and n1 < 0.5 * n0
