# dev: single-pass encoder validation (three encoder modes against the oracle), then short bench lines; A/B of the 64-bit literal table
run_all() {
for m in nocrc inline old; do timeout 250 python tests/gpu_fused_check.py $m 2>&1 | tail -14; done
for a in "c2 g1" "c3 g1" "c2 g0"; do set -- $a
timeout 200 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1 $2', d['value'], d['ms_per_step'], d.get('kernels_ms'), d['config'].get('parity_image0_vs_oracle'))"
done
}
echo "=== LIT64=1 (default build)"; run_all
echo "=== LIT64=0"; FPNGB_NVCC_DEFS=-DFPNGB_LIT64=0 python -c "
from fpng_b200 import _build; _build.build(force=True)" 2>&1 | tail -2
timeout 250 python tests/gpu_fused_check.py inline 2>&1 | tail -4
for a in "c2 g1" "c3 g1"; do set -- $a
timeout 200 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1 $2', d['value'], d['ms_per_step'], d.get('kernels_ms'), d['config'].get('parity_image0_vs_oracle'))"
done
echo "=== decode tests + bench decode leg (default build again)"
python -c "
from fpng_b200 import _build; _build.build(force=True)" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_decode_gpu.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('decode', d['decode']['value'], d['decode']['ms_per_step'], d['decode'].get('kernels_ms'), d['decode']['pixels_match_input'], d['decode']['input'][:60]); print('e2e', d['e2e']['value'])"
