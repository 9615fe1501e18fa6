"""world_size-2 gloo test of the multi-rank host logic used by bench.py: strided sharding of the candidate grid,
all_gather of the per-rank log-likelihood slices, and the log-sum-exp normalisation."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["GPCC_ROOT"])
from gpcc_b200.sharding import shard_indices, gather_strided
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
M = 101
rg = np.random.default_rng(0)
ll_all = rg.normal(-150.0, 3.0, M)
mine = shard_indices(M, rank, 2)
assert np.array_equal(mine, np.arange(rank, M, 2))
full = gather_strided(torch.from_numpy(ll_all[mine]), M, rank, 2).numpy()
assert np.array_equal(full, ll_all), (rank, np.abs(full - ll_all).max())
if rank == 0:
    print("OK")
dist.destroy_process_group()
'''


def test_strided_shard_and_allgather_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=port, GPCC_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "OK" in outs[0][0]


def test_shard_indices_cover_grid_exactly_once():
    sys.path.insert(0, ROOT)
    from gpcc_b200.sharding import shard_indices
    for M in (1, 7, 101, 10201):
        for W in (1, 2, 4, 8):
            allidx = np.concatenate([shard_indices(M, r, W) for r in range(W)])
            assert np.array_equal(np.sort(allidx), np.arange(M))
            sizes = [len(shard_indices(M, r, W)) for r in range(W)]
            assert max(sizes) - min(sizes) <= 1
