timeout 1800 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25
for a in "c2 g1" "c4 g1"; do set -- $a
timeout 200 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1 $2', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d.get('kernels_ms').items() if v}, d['config'].get('parity_image0_vs_oracle'))"
done
for enc in default fused; do
timeout 200 python bench.py --encoder $enc --workload odd --no-cpu --no-decode --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$enc odd', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d.get('kernels_ms').items() if v}, d['config'].get('parity_image0_vs_oracle'))"
done
