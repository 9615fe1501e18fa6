"""Per-step fit time of the cfg3 grid on one GPU, optionally with torch.distributed initialised (dev diagnostic)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import gpcc_b200
mode = os.environ.get("DIAG_MODE", "plain")
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
if mode in ("torch", "dist"):
    import torch
    torch.cuda.set_device(local)
    if mode == "dist":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        x = torch.ones(4, device="cuda"); dist.all_reduce(x); torch.cuda.synchronize()
ctx = gpcc_b200.Context(devices=[local], profiling=True)
t, y, s, d = gpcc_b200.simulatethreelightcurves()
p = gpcc_b200.Problem(t, y, s, "matern32", ctx)
c = np.arange(0.0, 20.0001, 0.2)
delays = np.array([[0.0, a, b] for b in c for a in c])
theta0 = gpcc_b200.initial_solutions(y, 1, 1, 5, 0.1, 300.0)[0][0]
ts = []
for i in range(10):
    t0 = time.perf_counter(); r = p.fit_batch(delays, theta0, iterations=1000, rhomin=0.1, rhomax=300.0); ts.append((time.perf_counter() - t0) * 1e3)
st = ctx.stats()
print("mode %s rank %d/%d: per-fit ms %s kernel_ms %.0f" % (mode, rank, world, ["%.0f" % v for v in ts], st["ms_eval_kernels"]), flush=True)
