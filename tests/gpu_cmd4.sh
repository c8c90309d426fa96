# A/B: pack LIT64 on (default build) vs off; odd (unaligned) workload: generic two-kernel vs single-pass direct-load; then the 2-GPU-free parts of the suite
for a in "c2 g1" "c3 g1" "c2 g0" "c4 g1"; do set -- $a
timeout 200 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('LIT64=1 $1 $2', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d.get('kernels_ms').items() if v}, d['config'].get('parity_image0_vs_oracle'))"
done
for enc in two_kernel_serial fused; do
timeout 200 python bench.py --encoder $enc --workload odd --no-cpu --no-decode --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$enc odd', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d.get('kernels_ms').items() if v}, d['config'].get('parity_image0_vs_oracle'))"
done
FPNGB_NVCC_DEFS=-DFPNGB_PACK_LIT64=0 python -c "
from fpng_b200 import _build; _build.build(force=True)" 2>&1 | tail -1
for a in "c2 g1" "c3 g1"; do set -- $a
timeout 200 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('LIT64=0 $1 $2', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d.get('kernels_ms').items() if v}, d['config'].get('parity_image0_vs_oracle'))"
done
