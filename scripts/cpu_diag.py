import os, time, multiprocessing as mp
def burn(_):
    t0 = time.perf_counter(); x = 0
    for i in range(20_000_000): x += i * i
    return time.perf_counter() - t0
if __name__ == "__main__":
    print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/sys/fs/cgroup/cpu.stat"):
        try: print(f, open(f).read().strip().replace("\n", " | "))
        except Exception as e: print(f, "n/a")
    for n in (1, 2, 4, 8, 16):
        with mp.Pool(n) as p:
            t0 = time.perf_counter(); r = p.map(burn, range(n)); dt = time.perf_counter() - t0
        print(n, "procs: wall %.2f s, per-proc %.2f..%.2f" % (dt, min(r), max(r)))
