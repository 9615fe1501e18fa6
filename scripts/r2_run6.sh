#!/bin/bash
# round 2: 2 GPUs: NM tests, multi-device tests, bench N=2
cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 180 > gpurun_out/r2_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multi.log
tail -5 gpurun_out/r2_pytest_multi.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 180 -s -k "nelder or prior_with or alternative" > gpurun_out/r2_pytest_nm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_nm.log
grep -v "^Running\|^\s*$\|iterations \|initialrandom\|numberofrestarts\|JITTER\|ρm\|Σb\|Overall\|unpack\|Initial\|^\s[0-9.]*$" gpurun_out/r2_pytest_nm.log | tail -12
