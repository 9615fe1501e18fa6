#!/bin/bash
# round 2: ncu launch list of the bench command + one full capture of the persistent fit kernel
cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline"
$CMD > gpurun_out/r2_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/r2_ncu_launch.log 2>&1
tail -3 gpurun_out/r2_ncu_launch.log
python scripts/prof_fit.py > gpurun_out/r2_plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:small_fit -s 1 -c 1 -o gpurun_out/prof_small_fit_r2 python scripts/prof_fit.py > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log
ls -la gpurun_out/*.ncu-rep | tail -2
