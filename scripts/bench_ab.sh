cd /root/repo
# usage: CFGS="one_stream:merge_below ..." bash scripts/bench_ab.sh
for o in ${CFGS:-0:1024 1:1024}; do os=${o%%:*}; mb=${o##*:}; GPCC_ONE_STREAM=$os GPCC_MERGE_BELOW=$mb timeout 250 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('one_stream=$os merge_below=$mb', round(d['value']), 'cand/s', round(d['ms_per_step'],1), 'ms/step; kernel ms', round(d['roofline']['kernel_ms_per_step'],1), 'TF', round(d['roofline']['achieved'],2), 'e2e', round(d['e2e']['value']), 'cfg2 ms', round(d['also']['cfg2']['ms_per_grid'],2), 'launches', d['gpu_launches'])"; done
