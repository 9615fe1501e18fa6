#!/bin/bash
# round 2, GPU call 1: parity suite + small-kernel variants
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -15 gpurun_out/r2_pytest1.log
( timeout 200 python scripts/time_small.py
  GPCC_SMALL_NO_FWD=1 timeout 200 python scripts/time_small.py
  GPCC_SMALL_VARIANT=6 timeout 200 python scripts/time_small.py
  GPCC_SMALL_VARIANT=0 timeout 200 python scripts/time_small.py ) > gpurun_out/r2_time_small1.log 2>&1
cat gpurun_out/r2_time_small1.log
