#!/bin/bash
# round 2: lean step + library-owned collective: parity suite, bench N=1
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -6 gpurun_out/r2_pytest4.log
GPCC_FIT_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; tail -2 gpurun_out/r2_bench4.err; cat gpurun_out/r2_bench4.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench4_ref.json 2> gpurun_out/r2_bench4_ref.err; cat gpurun_out/r2_bench4_ref.json
