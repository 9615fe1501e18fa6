#!/bin/bash
# round 2, GPU call: small-kernel timing (forward-only mode on/off) + parity suite
cd /root/repo
mkdir -p gpurun_out
( timeout 90 python scripts/time_small.py
  GPCC_SMALL_NO_FWD=1 timeout 90 python scripts/time_small.py ) > gpurun_out/r2_time_small2.log 2>&1
cat gpurun_out/r2_time_small2.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -40 gpurun_out/r2_pytest2.log
