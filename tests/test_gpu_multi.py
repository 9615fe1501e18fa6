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


RANK_WORKER = r'''
import os, sys, time
import numpy as np
sys.path.insert(0, os.environ["GPCC_ROOT"]); sys.path.insert(0, os.path.join(os.environ["GPCC_ROOT"], "tests"))
import gpcc_b200
from conftest import load_golden
rank, world, idfile = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), os.environ["IDFILE"]
ctx = gpcc_b200.Context(devices=[rank])
if rank == 0:
    with open(idfile + ".tmp", "wb") as f:
        f.write(gpcc_b200.comm_unique_id())
    os.rename(idfile + ".tmp", idfile)
while not os.path.exists(idfile):
    time.sleep(0.05)
ctx.comm_init_rank(world, rank, open(idfile, "rb").read())
g = load_golden("fit_cfg1_cfg2")
delays = np.stack([np.zeros_like(g["cands"]), g["cands"]], 1)[:77]          # not a multiple of the world size
lp = gpcc_b200.uniformpriordelay(L=1e44, z=0.0).logpdf(g["cands"])[:77]
p = gpcc_b200.Problem(g["tb"], g["yb"], g["sb"], "matern32", ctx)
r = p.grid_posterior(delays, g["theta0"], iterations=1000, rhomin=0.1, rhomax=300.0, logprior=lp)
np.savez(os.environ["OUTFILE"] + str(rank), **r)
'''


@pytest.mark.skipif("_ndev() < 2")
def test_one_process_per_gpu_library_allgather(ctx, tmp_path):
    """The torchrun layout (bench.py): one process per GPU, each with a one-device context joined through
    gpcc_ctx_comm_init_rank; gpcc_grid_posterior shards the grid (candidate m on rank m mod world), runs the library's own
    ncclAllGather of the per-candidate records and returns the FULL outputs on every rank -- identical to one device."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "rank_worker.py"
    script.write_text(RANK_WORKER)
    env = dict(os.environ, GPCC_ROOT=root, WORLD_SIZE="2", IDFILE=str(tmp_path / "nccl_id"), OUTFILE=str(tmp_path / "out"))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [q.communicate(timeout=150) for q in procs]
    for q, (o, e) in zip(procs, outs):
        assert q.returncode == 0, e[-3000:]
    g = load_golden("fit_cfg1_cfg2")
    delays = np.stack([np.zeros_like(g["cands"]), g["cands"]], 1)[:77]
    lp = gpcc_b200.uniformpriordelay(L=1e44, z=0.0).logpdf(g["cands"])[:77]
    p1 = Problem(g["tb"], g["yb"], g["sb"], "matern32", ctx)
    r1 = p1.grid_posterior(delays, g["theta0"], iterations=1000, rhomin=0.1, rhomax=300.0, logprior=lp)
    from gpcc_b200.sharding import shard_indices
    for rank in range(2):
        z = np.load(str(tmp_path / "out") + str(rank) + ".npz")
        for key in ("loglikel", "theta", "alpha", "rho", "nfev", "info"):
            assert np.array_equal(z[key], r1[key]), (rank, key)           # same kernels, same inputs: bitwise, on every rank
        assert np.max(np.abs(z["posterior"] - r1["posterior"])) < 1e-14 and abs(z["posterior"].sum() - 1.0) < 1e-12
    assert np.array_equal(shard_indices(77, 1, 2), np.arange(1, 77, 2))     # the layout the library uses (gpcc_b200/sharding.py)
