cd /root/repo
for w in ${WAVES:-1 64}; do for h in ${HELPERS:-0 1}; do GPCC_FRAG_HELPER=$h GPCC_TS_WAVES=$w GPCC_B200_LIB=/root/repo/gpcc_b200/libgpcc_b200_prof.so GPCC_SMALL_FRAG=1 timeout ${TMO:-60} python scripts/time_small.py 2>&1 | grep "frag tl\|N=1\|rror\|small_frag" | grep -v "T=14" | tail -${TAILN:-10} | sed "s/^/[waves=$w helper=$h] /"; echo "[waves=$w helper=$h] rc=${PIPESTATUS[0]}"; done; done > gpurun_out/prof_frag.log 2>&1
cat gpurun_out/prof_frag.log
