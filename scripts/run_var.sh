cd /root/repo
for w in ${WAVES:-64}; do for v in ${VARS:-1 3}; do GPCC_TS_WAVES=$w GPCC_SMALL_VARIANT=$v timeout 120 python scripts/time_small.py 2>&1 | grep "N=1\|rror" | sed "s/^/[waves=$w variant=$v] /"; done; done > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
