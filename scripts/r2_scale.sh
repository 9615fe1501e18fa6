#!/bin/bash
# round 2: bench at N GPUs (strong scaling default; weak + cfg4 under "also")
N=$1
cd /root/repo
mkdir -p gpurun_out
GPCC_FIT_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/final7_bench_n$N.json 2> gpurun_out/final7_bench_n$N.err
echo "rc=$?"; grep "gpcc fit" gpurun_out/final7_bench_n$N.err | tail -3; tail -2 gpurun_out/final7_bench_n$N.err | cut -c1-300; cat gpurun_out/final7_bench_n$N.json
