# dev: single-pass encoder validation (three encoder modes against the oracle), then a short bench line
for m in nocrc inline old; do timeout 250 python tests/gpu_fused_check.py $m 2>&1 | tail -14; done
for a in "c2 g1" "c3 g1" "c2 g0"; do set -- $a
timeout 200 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1 $2', d['value'], d['ms_per_step'], d.get('kernels_ms'), d['config'].get('parity_image0_vs_oracle'))"
done
