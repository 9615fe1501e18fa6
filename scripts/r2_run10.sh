#!/bin/bash
# round 2: state after kernel work: full parity suite + full bench at N=1 (with cpu baseline and "also")
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -4 gpurun_out/r2_pytest6.log
GPCC_FIT_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; grep "gpcc fit" gpurun_out/r2_bench_n1.err | sed -n '5p'; cat gpurun_out/r2_bench_n1.json
