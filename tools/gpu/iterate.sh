# quick GPU iteration: encode parity tests, then short bench lines (no CPU arm) for the four main workloads
timeout 400 python -m pytest tests/test_encode_gpu.py -x -q 2>&1 | tail -4
for a in "c2 g1" "c3 g1" "c2 g0" "c4 g1"; do set -- $a
  timeout 200 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1 $2', d['value'], d['ms_per_step'], d.get('kernels_ms'), d.get('parity_image0_vs_oracle'))"
done
