#!/bin/bash
# round 2 final evidence, 1 GPU: full parity suite, smoke, bench (driver flags), reference arm, ncu launch list + full capture
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log; tail -3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/final_bench_n1.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cat gpurun_out/final_bench_ref.json | cut -c1-400
CMD="python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2_final.csv $CMD > gpurun_out/final_ncu_launch.log 2>&1; tail -1 gpurun_out/final_ncu_launch.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:small_fit -s 1 -c 1 -o gpurun_out/prof_small_fit_r2_final python scripts/prof_fit.py > gpurun_out/final_ncu_full.log 2>&1; tail -2 gpurun_out/final_ncu_full.log
