cd /root/repo
for o in ${OPTS:-"0 1024" "0 256" "0 64" "1 1024"}; do set -- $o; GPCC_ONE_STREAM=$1 GPCC_MERGE_BELOW=$2 timeout 250 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('one_stream=$1 merge_below=$2', round(d['value']), 'cand/s', round(d['ms_per_step'],1), 'ms/step; kernel ms', round(d['roofline']['kernel_ms_per_step'],1), 'TF', round(d['roofline']['achieved'],2), 'e2e', round(d['e2e']['value']), 'cfg2 ms', round(d['also']['cfg2']['ms_per_grid'],2), 'launches', d['gpu_launches'])"; done
