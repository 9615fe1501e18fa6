"""Multi-GPU tests (need >= 2 B200s; run with gpurun --gpus 2 -- python -m pytest tests -m gpu):
the grid sharded over the devices of ONE context (candidate m -> device m mod ndev) with a single NCCL allgather
of the per-candidate log-likelihoods gives the same numbers as one device."""
import numpy as np
import pytest

import gpcc_b200
from conftest import load_golden
from gpcc_b200 import Context, Problem

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_ndev() < 2")
def test_sharded_grid_posterior_matches_single_device(ctx):
    g = load_golden("fit_cfg1_cfg2")
    delays = np.stack([np.zeros_like(g["cands"]), g["cands"]], 1)
    p1 = Problem(g["tb"], g["yb"], g["sb"], "matern32", ctx)
    r1 = p1.grid_posterior(delays, g["theta0"], iterations=1000, rhomin=0.1, rhomax=300.0)
    ctx2 = Context(2)
    assert ctx2.ndev == 2
    p2 = Problem(g["tb"], g["yb"], g["sb"], "matern32", ctx2)
    r2 = p2.grid_posterior(delays, g["theta0"], iterations=1000, rhomin=0.1, rhomax=300.0)
    assert np.array_equal(r1["loglikel"], r2["loglikel"])                 # same kernels, same inputs: bitwise
    assert np.max(np.abs(r1["posterior"] - r2["posterior"])) < 1e-14      # NCCL allgather + log-sum-exp
    assert r2["posterior"].sum() == pytest.approx(1.0, abs=1e-12)
    assert np.max(np.abs(r2["posterior"] - g["post_flat"])) < 1e-4
    lp = gpcc_b200.uniformpriordelay(L=1e44, z=0.0).logpdf(g["cands"])
    r3 = p2.grid_posterior(delays[:77], g["theta0"], iterations=1000, rhomin=0.1, rhomax=300.0, logprior=lp[:77])   # M not a multiple of ndev
    r4 = p1.grid_posterior(delays[:77], g["theta0"], iterations=1000, rhomin=0.1, rhomax=300.0, logprior=lp[:77])
    assert np.max(np.abs(r3["posterior"] - r4["posterior"])) < 1e-14
    ll, info = p2.loglik_batch(delays, r1["alpha"], r1["rho"])
    assert np.allclose(ll, r1["loglikel"], rtol=1e-12)
