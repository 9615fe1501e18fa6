"""One-process-per-GPU sharding of the candidate grid (README.md:202,285 `pmap`).

Candidates are independent, so rank r fits the strided slice m = r, r+W, r+2W, ... (balances the smoothly
varying iteration counts over the grid) with no communication, and the only collective on the path is one
all-gather of the per-candidate log-likelihoods (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_indices(M, rank, world):
    return np.arange(rank, M, world)


def gather_strided(local, M, rank, world):
    """all_gather of the strided slices back into candidate order.  `local` is a 1-D float64 torch tensor on the
    rank's device holding the values of shard_indices(M, rank, world)."""
    import torch
    import torch.distributed as dist
    per = (M + world - 1) // world
    buf = torch.full((per,), float("-inf"), dtype=torch.float64, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty((world, per), dtype=torch.float64, device=local.device)
    if world > 1:
        dist.all_gather_into_tensor(out.view(-1), buf)
    else:
        out[0] = buf
    return out.t().reshape(-1)[:M].contiguous()      # [k][rank] -> m = rank + k*world
