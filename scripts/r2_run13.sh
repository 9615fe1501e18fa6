#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 8 --single-process --steps 10 --warmup 3 > gpurun_out/r2_bench_sp8.json 2> gpurun_out/r2_bench_sp8.err; echo rc=$?; tail -2 gpurun_out/r2_bench_sp8.err | cut -c1-300; cat gpurun_out/r2_bench_sp8.json
timeout 600 python bench.py --gpus 8 --single-process --scaling weak --steps 5 --warmup 3 > gpurun_out/r2_bench_sp8w.json 2> gpurun_out/r2_bench_sp8w.err; cat gpurun_out/r2_bench_sp8w.json | cut -c1-250
bash scripts/r2_scale.sh 8 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); a=d['also']; print('torchrun strong', d['value'], d['ms_per_step'], 'weak', a['scaling_weak']['candidates_per_s'], 'cfg4', a['cfg4']['seconds'], a['cfg4']['structure_reuse'], a['cfg4']['cholesky_frac_of_fp64_peak'])"
