#!/bin/bash
# round 2, GPU call: device-resident fit: parity suite + bench A/B against the host-driven loop
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -s > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
grep -v "^Running\|^\s*$\|iterations \|initialrandom\|numberofrestarts\|JITTER\|ρm\|Σb\|Overall\|unpack\|Initial\|^\s[0-9.]*$" gpurun_out/r2_pytest3.log | tail -30
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2_bench_dev.json 2> gpurun_out/r2_bench_dev.err; tail -3 gpurun_out/r2_bench_dev.err; cat gpurun_out/r2_bench_dev.json
GPCC_FIT_HOST=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2_bench_host.json 2> gpurun_out/r2_bench_host.err; cat gpurun_out/r2_bench_host.json
GPCC_SCREEN_FULL=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2_bench_screenfull.json 2> gpurun_out/r2_bench_screenfull.err; cat gpurun_out/r2_bench_screenfull.json
