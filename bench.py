#!/usr/bin/env python
"""bench.py -- delay candidates/sec (fitted logL+grad) of the B200-native GPCC hot path.

A "step" is one pass of the hot path over one batch of synthetic input: fit every candidate of a delay grid
(5 screening evaluations + batched L-BFGS on the analytic gradient, all likelihood work in CUDA), all-gather the
per-candidate log-likelihoods across ranks (NCCL) and normalise them into the posterior under the delay prior.
Default workload = BASELINE.json configs[2] (simulatethreelightcurves-style data, 2-D grid (0:0.2:20)^2 = 10 201
candidates per GPU, matern32, rhomax=300, iterations=1000) -- the configuration the metric "candidates/sec at
1/2/4/8 B200" is quoted on; it is scaled weakly (rank r gets its own 10 201-candidate slice of a grid refined
N-fold along tau_3).  configs[1] (101 candidates) cannot occupy 148 SMs; it is reported under "also".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RHOMIN, RHOMAX, ITERATIONS, INITIALRANDOM = 0.1, 300.0, 1000, 5
METRIC = "delay candidates/sec (fitted logL+grad)"


def make_workload(name, world):
    """Returns (tarray, yarray, sarray, delays[M_total][L], label).  Data generator: gpcc_b200.synthetic (host, numpy)."""
    from gpcc_b200 import synthetic
    if name == "cfg3":
        t, y, s, _ = synthetic.simulatethreelightcurves()
        c2 = np.arange(0.0, 20.0001, 0.2)
        c3 = np.linspace(0.0, 20.0, 101 * world)
        delays = np.array([[0.0, a, b] for b in c3 for a in c2])          # d1 fastest (README.md:231-235)
        label = "cfg3: simulatethreelightcurves-style N=(60,50,40), grid (0:0.2:20) x %d points in [0,20] = %d candidates (%d per GPU), matern32" % (len(c3), len(delays), len(delays) // world)
    elif name == "cfg2":
        t, y, s, _ = synthetic.simulatetwolightcurves()
        c = np.linspace(0.0, 10.0, 101 * world) if world > 1 else np.arange(0.0, 10.0001, 0.1)
        delays = np.stack([np.zeros_like(c), c], 1)
        label = "cfg2: simulatetwolightcurves-style N=(60,50), 1-D grid over [0,10] = %d candidates, matern32" % len(delays)
    else:
        raise SystemExit("unknown workload " + name)
    return t, y, s, delays, label


# ---- CPU baseline: the oracle (restated reference, Nelder-Mead like the reference) over a process pool -------------
_W = {}


def _cpu_init(tys, theta0):
    try:
        from threadpoolctl import threadpool_limits
        _W["lim"] = threadpool_limits(1)            # OpenBLAS pinned to 1 thread per worker
    except Exception:
        pass
    _W["tys"], _W["theta0"] = tys, theta0


def _cpu_fit(dl):
    import oracle
    t, y, s = _W["tys"]
    return oracle.gpcc(t, y, s, kernel="matern32", delays=dl, iterations=ITERATIONS, rhomin=RHOMIN, rhomax=RHOMAX,
                       theta0=_W["theta0"][None], optimizer="neldermead")[0]


def cpu_reference_rate(t, y, s, delays, theta0, n_sample, steps=1, warmup=0):
    """README `pmap` recipe on the host cores with the oracle: one candidate per task, workers = all cores."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    idx = np.unique(np.linspace(0, len(delays) - 1, n_sample).astype(int))
    sample = delays[idx]
    ctxmp = mp.get_context("fork")
    with ctxmp.Pool(cores, initializer=_cpu_init, initargs=((t, y, s), theta0)) as pool:
        for _ in range(warmup):
            pool.map(_cpu_fit, sample[:cores], chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            ll = pool.map(_cpu_fit, sample, chunksize=1)
        dt = (time.perf_counter() - t0) / steps
    return len(sample) / dt, cores, len(sample), dt, np.array(ll), idx


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world, out):
    """--impl reference: the reference's own CPU path (restated: no Julia in this image) on the host cores."""
    if rank != 0:
        return
    t, y, s, delays, label = make_workload(args.workload, 1)
    import gpcc_b200
    theta0 = gpcc_b200.initial_solutions(y, 1, 1, INITIALRANDOM, RHOMIN, RHOMAX)[0][0]
    cores = os.cpu_count() or 1
    n_sample = args.cpu_sample or 2 * cores
    rate, cores, ns, dt, _, _ = cpu_reference_rate(t, y, s, delays, theta0, n_sample, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "optimizer": "Nelder-Mead g_tol=1e-6 (the reference's), iterations=1000, initialrandom=5",
                       "note": "restated reference (numpy/scipy oracle), not Julia: no julia toolchain in this image"},
            "cpu_baseline": {"value": rate, "unit": "candidates/s", "cores": cores, "kind": "port",
                             "sample": "%d of %d candidates (evenly spaced over the grid) per step, one candidate per task, 1 BLAS thread per worker" % (ns, len(delays))},
            "e2e": {"value": rate, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner) write to fd 1, so fd 1 is
    # pointed at stderr for the duration of the run and the JSON line is written to the saved descriptor.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg2"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, real_stdout)
        return
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N>1 launch with python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    warmup = max(args.warmup, 3)

    t, y, s, delays_all, label = make_workload(args.workload, world)
    import gpcc_b200
    theta0 = gpcc_b200.initial_solutions(y, 1, 1, INITIALRANDOM, RHOMIN, RHOMAX)[0][0]

    # ---- CPU baseline first (fork before any CUDA state exists), rank 0 at N=1 only --------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, cores, ns, dt, ll_cpu, idx_cpu = cpu_reference_rate(t, y, s, delays_all, theta0, args.cpu_sample or 4 * cores)
        cpu = {"value": rate, "unit": "candidates/s", "cores": cores, "kind": "port",
               "sample": "%d of %d candidates (evenly spaced), oracle = restated reference with Nelder-Mead, pool of %d processes x 1 BLAS thread, %.1f s" % (ns, len(delays_all), cores, dt)}

    import torch
    import torch.distributed as dist
    from gpcc_b200.sharding import shard_indices, gather_strided
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = gpcc_b200.Context(devices=[local_rank], profiling=True)
    mine = shard_indices(len(delays_all), rank, world)
    delays = np.ascontiguousarray(delays_all[mine])
    M_total, M_local = len(delays_all), len(delays)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    problem = gpcc_b200.Problem(t, y, s, gpcc_b200.matern32, ctx)
    state = {}

    def step_resident():
        res = problem.fit_batch(delays, theta0, iterations=ITERATIONS, rhomin=RHOMIN, rhomax=RHOMAX)
        full = gather_strided(torch.from_numpy(res["loglikel"]).to(dev), M_total, rank, world)     # NCCL all-gather
        post = ctx.getprobabilities(full.cpu().numpy())                                        # log-sum-exp on device
        state.update(res=res, post=post, stats=ctx.stats())
        return post

    def step_e2e():
        # the user-facing call with HOST buffers: upload the light curves, fit the grid, posterior back on the host
        pr = gpcc_b200.Problem(t, y, s, gpcc_b200.matern32, ctx)
        res = pr.fit_batch(delays, theta0, iterations=ITERATIONS, rhomin=RHOMIN, rhomax=RHOMAX)
        full = gather_strided(torch.from_numpy(res["loglikel"]).to(dev), M_total, rank, world)
        post = ctx.getprobabilities(full.cpu().numpy())
        state.update(stats_e2e=ctx.stats(), res_e2e=res)
        pr.close()
        return post

    def timed(fn, steps):
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        wall = 0.0
        for i in range(steps):
            flush.fill_(float(i))              # L2 flush between timed iterations (untimed)
            barrier()
            e0[i].record()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            e1[i].record()
            wall += time.perf_counter() - t0
            barrier()
        if os.environ.get("GPCC_BENCH_VERBOSE"):
            sys.stderr.write("rank %d %s per-step ms: %s\n" % (rank, fn.__name__, ["%.0f" % a.elapsed_time(b) for a, b in zip(e0, e1)]))
        dev_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
        tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item()) / steps, wall * 1e3 / steps

    for _ in range(warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("GPCC_BENCH_NO_SAMPLER"):
        sampler.start()
    # count launches / kernel time per timed step through the library's own CUDA-event statistics
    kstats = []

    phase_ms = []

    def step_resident_counted():
        t0 = time.perf_counter()
        res = problem.fit_batch(delays, theta0, iterations=ITERATIONS, rhomin=RHOMIN, rhomax=RHOMAX)
        t1 = time.perf_counter()
        kstats.append(ctx.stats())
        full = gather_strided(torch.from_numpy(res["loglikel"]).to(dev), M_total, rank, world)
        host_ll = full.cpu().numpy()
        t2 = time.perf_counter()
        post = ctx.getprobabilities(host_ll)
        t3 = time.perf_counter()
        phase_ms.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, kstats[-1]["ms_eval_kernels"]))
        state.update(res=res, post=post)
        return post

    ms_step, wall_ms = timed(step_resident_counted, args.steps)
    if os.environ.get("GPCC_BENCH_VERBOSE"):
        sys.stderr.write("rank %d phases (fit, allgather, posterior, fit-kernel) ms: %s\n" % (rank, [tuple(round(v) for v in p_) for p_ in phase_ms]))
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (fused small-N sweep), from the library's CUDA events on its own stream ----
    N = int(sum(len(a) for a in t))
    L = len(t)
    evals = sum(k["n_evals_grad"] for k in kstats)
    kern_ms = sum(k["ms_eval_kernels"] for k in kstats)
    n_launch = sum(k["n_eval_launches"] for k in kstats) + args.steps       # + 1 posterior kernel per step
    peaks = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r1.json")))
    achieved = evals * float(N) ** 3 / (kern_ms * 1e-3) / 1e12               # algorithmic flops: N^3 per logL+grad evaluation
    traffic = None
    tp = os.path.join(ROOT, "profiles", "small_sweep_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "pipe": "FP64 (DFMA pipe; the DMMA peak is the same on B200: 37.0 vs 36.7 TFLOP/s measured)",
                "kernel": "small_sweep_kernel (fused assembly + symmetric sweep + gradient, one CTA per evaluation)",
                "achieved": achieved, "peak": peaks["dfma_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["dfma_tflops"],
                "traffic": traffic, "peak_source": "measured, profiles/fp64_peaks_r1.json (MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flops_per_eval": float(N) ** 3, "evals_per_step": evals / args.steps,
                "kernel_ms_per_step": kern_ms / args.steps, "kernel_share_of_step": kern_ms / args.steps / ms_step}

    nfev = state["res"]["nfev"]
    per_eval_h2d = (2 * L + 1) * 8
    per_eval_d2h = (L + 2) * 8 + 4
    data_bytes = N * (4 * 8 + 4)
    e2e_evals = state["stats_e2e"]["n_evals"]
    line = {"metric": METRIC, "value": M_total / (ms_step * 1e-3), "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "kernel": "matern32", "rhomin": RHOMIN, "rhomax": RHOMAX, "iterations": ITERATIONS,
                       "initialrandom": INITIALRANDOM, "optimizer": "host-driven batched L-BFGS (m=8), analytic gradient on device",
                       "candidates_per_gpu": M_local, "l2": "256 MB flush buffer written between timed steps",
                       "parallelism": "candidate grid sharded strided over %d rank(s); one NCCL all-gather of log-likelihoods" % world,
                       "mean_nfev": float(np.mean(nfev)), "max_nfev": int(np.max(nfev))},
            "e2e": {"value": M_total / (ms_e2e * 1e-3), "unit": "candidates/s",
                    "h2d_bytes_per_step": int(data_bytes + e2e_evals * per_eval_h2d + M_total * 8),
                    "d2h_bytes_per_step": int(e2e_evals * per_eval_d2h + M_total * 8), "ms_per_step": ms_e2e},
            "gpu_launches": int(n_launch), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "wall_ms_per_step": wall_ms}
    if rank == 0 and world == 1 and not args.no_also and args.workload == "cfg3":
        # configs[1] (101 candidates, 2 bands) for the record
        t2, y2, s2, d2, label2 = make_workload("cfg2", 1)
        p2 = gpcc_b200.Problem(t2, y2, s2, gpcc_b200.matern32, ctx)
        th2 = gpcc_b200.initial_solutions(y2, 1, 1, INITIALRANDOM, RHOMIN, RHOMAX)[0][0]
        for _ in range(3):
            p2.grid_posterior(d2, th2, iterations=ITERATIONS, rhomin=RHOMIN, rhomax=RHOMAX)
        t0 = time.perf_counter()
        for _ in range(5):
            p2.grid_posterior(d2, th2, iterations=ITERATIONS, rhomin=RHOMIN, rhomax=RHOMAX)
        dt = (time.perf_counter() - t0) / 5
        line["also"] = {"cfg2": {"workload": label2, "candidates_per_s": len(d2) / dt, "ms_per_grid": dt * 1e3}}
        # configs[3] (3 bands x 2048 points, N=6144, matern52): the tiled large-N path.  A fitted grid at this size is
        # hours of work, so the record is the fixed-hyper-parameter sweep rate (one factorisation per candidate, the
        # workload BASELINE.md section 3 ties the "<10 s on 8 GPUs" target to) and the logL+gradient rate.
        try:
            from gpcc_b200 import synthetic
            t4, y4, s4, _ = synthetic.synthetic_bands([2048, 2048, 2048], seed=4)
            p4 = gpcc_b200.Problem(t4, y4, s4, gpcc_b200.matern52, ctx)
            N4, M4 = 6144, 16
            rg = np.random.default_rng(2)
            d4 = np.zeros((M4, 3)); d4[:, 1:] = rg.uniform(0.0, 19.8, (M4, 2))
            a4, r4 = np.tile([1.0, 2.2, 4.0], (M4, 1)), np.full(M4, 3.5)
            out4 = {}
            for grad in (False, True):
                for _ in range(2):
                    p4.loglik_batch(d4, a4, r4, want_grad=grad)
                t0 = time.perf_counter()
                p4.loglik_batch(d4, a4, r4, want_grad=grad)
                dt4 = time.perf_counter() - t0
                st4 = ctx.stats()
                fl = M4 * float(N4) ** 3 * (1.0 if grad else 1.0 / 3.0)
                key = "logL+grad (symmetric sweep, N^3 flop)" if grad else "logL only (blocked Cholesky, N^3/3 flop)"
                out4[key] = {"evaluations_per_s": M4 / dt4, "ms_per_evaluation": dt4 * 1e3 / M4,
                             "factor_tflops": fl / (st4["ms_factor"] * 1e-3) / 1e12,
                             "factor_frac_of_fp64_peak": fl / (st4["ms_factor"] * 1e-3) / 1e12 / peaks["dmma_m8n8k4_tflops"],
                             "assembly_GBps": M4 * 4.0 * N4 * (N4 + 1) / (st4["ms_assembly"] * 1e-3) / 1e9}
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
            ch = out4["logL only (blocked Cholesky, N^3/3 flop)"]
            line["also"]["cfg4"] = {"workload": "3 bands x 2048 points (N=6144), matern52, %d candidates per batch, fixed hyper-parameters" % M4,
                                    "results": out4, "assembly_frac_of_hbm_peak": ch["assembly_GBps"] / hbm,
                                    "hbm_peak_GBps": hbm, "fp64_peak_tflops": peaks["dmma_m8n8k4_tflops"],
                                    "projected_s_for_1e4_candidates_on_8_gpus_logL_only": 1e4 / 8 / ch["evaluations_per_s"]}
            p4.close()
        except Exception as e:      # never lose the headline line over the side measurement
            line["also"]["cfg4"] = {"error": repr(e)}
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
