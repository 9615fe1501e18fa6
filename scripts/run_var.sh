cd /root/repo
for v in 1 3; do GPCC_SMALL_VARIANT=$v timeout 120 python scripts/time_small.py 2>&1 | grep "N=1\|rror" | sed "s/^/[variant=$v] /"; done > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
