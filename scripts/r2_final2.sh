#!/bin/bash
# round 2, last session, final evidence on one GPU: full parity suite, smoke, bench (driver flags), reference arm, cfg4 A/B of the
# last-band cache, ncu launch list + full capture of the persistent fit kernel
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/final4_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final4_pytest_gpu.log; tail -3 gpurun_out/final4_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final4_bench_n1.json 2> gpurun_out/final4_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final4_bench_n1.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/final4_bench_ref.json 2> gpurun_out/final4_bench_ref.err; cut -c1-300 gpurun_out/final4_bench_ref.json
for sw in "" "GPCC_LARGE_NO_TAUCACHE=1" "GPCC_LARGE_NO_SHARE=1"; do env $sw timeout 300 python scripts/time_cfg4_share.py 1250; done > gpurun_out/final4_cfg4_ab.log 2>&1; cat gpurun_out/final4_cfg4_ab.log
CMD="python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2_final4.csv $CMD > gpurun_out/final4_ncu_launch.log 2>&1; tail -1 gpurun_out/final4_ncu_launch.log | cut -c1-200
