#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
GPCC_FIT_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; grep "gpcc fit" gpurun_out/r2_bench7.err | tail -1
python -c "
import json; d=json.load(open('gpurun_out/r2_bench7.json')); print('ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'kernel ms', d['roofline']['kernel_ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 -k "nelder or cfg2 or cfg1 or golden or device_resident" 2>&1 | tail -3
