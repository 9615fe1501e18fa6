#!/bin/bash
# round 2: 2 GPUs: multi-device tests (one process / two devices, and one process per device) + bench N=2
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 300 > gpurun_out/r2_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multi.log
tail -15 gpurun_out/r2_pytest_multi.log
GPCC_FIT_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; grep "gpcc fit" gpurun_out/r2_bench_n2.err | tail -2; tail -3 gpurun_out/r2_bench_n2.err; cat gpurun_out/r2_bench_n2.json
GPCC_FIT_DEBUG=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err; grep "gpcc fit" gpurun_out/r2_bench_n1b.err | tail -1; cat gpurun_out/r2_bench_n1b.json | head -c 400
