#!/bin/bash
# round 2: rolled assembly/gradient (instruction footprint), parallel unpack: parity suite, microbench, bench
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -4 gpurun_out/r2_pytest5.log
(cd scripts/microbench && timeout 60 ./step_time | grep "M= 2368") 
timeout 100 python scripts/time_small.py 2>&1 | tail -4
GPCC_FIT_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; grep "gpcc fit" gpurun_out/r2_bench5.err | tail -1; cat gpurun_out/r2_bench5.json | head -c 330; echo
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_bench5.json')); print(d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'])
PY
