set -x
cd /root/repo
./scripts/microbench/fp64_latency > gpurun_out/fp64_latency.log 2>&1
for v in "0 0" "1 0" "1 1"; do set -- $v; GPCC_SMALL_BLOCK=$1 GPCC_BLOCK_VARIANT=$2 timeout 300 python scripts/time_small.py 2>&1 | sed "s/^/[block=$1 var=$2] /"; done > gpurun_out/time_block.log 2>&1
GPCC_SMALL_BLOCK=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_block.log 2>&1
tail -5 gpurun_out/pytest_block.log
cat gpurun_out/time_block.log gpurun_out/fp64_latency.log
