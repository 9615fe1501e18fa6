#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
GPCC_FIT_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2_bench_dev2.json 2> gpurun_out/r2_bench_dev2.err; grep "gpcc fit" gpurun_out/r2_bench_dev2.err | tail -4; cat gpurun_out/r2_bench_dev2.json | head -c 600
