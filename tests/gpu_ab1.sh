# round-2 session-2 A/B run 1: scan micro-optimisations + uniform warp index + unfilter unaligned stores
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/s2a_pytest_gpu.txt; cat gpurun_out/s2a_pytest_gpu.txt
for a in "c2 g1" "c3 g1" "odd g1" "c2 g0"; do set -- $a; timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --steps 10 > gpurun_out/s2a_bench_$1_$2.json 2>> gpurun_out/s2a_err.log; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s2a_bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value']), d['ms_per_step'], {k:round(v,3) for k,v in d['kernels_ms'].items() if v}, 'dec', round(d.get('decode',{}).get('value',0)), {k:round(v,3) for k,v in d.get('decode',{}).get('kernels_ms',{}).items()})
P
