for m in inline old old_serial; do timeout 250 python tests/gpu_fused_check.py $m 2>&1 | tail -3; done
for enc in two_kernel two_kernel_serial fused; do
for a in "c2 g1" "c3 g1" "c2 g0"; do set -- $a
timeout 200 python bench.py --encoder $enc --workload $1 --kind $2 --no-cpu --no-decode --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$enc $1 $2', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d.get('kernels_ms').items() if v}, d['config'].get('parity_image0_vs_oracle'))"
done; done
