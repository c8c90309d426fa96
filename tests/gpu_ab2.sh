# session-2 run 2: staged decode write kernel
set -x
timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_reference_fixtures_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2b_pytest.txt; cat gpurun_out/s2b_pytest.txt
for a in "c2 g1" "c3 g1" "c2 g0" "c4 g1" "odd g1"; do set -- $a; timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --steps 10 > gpurun_out/s2b_bench_$1_$2.json 2>> gpurun_out/s2b_err.log; done
FPNGB_DEC_STAGED=0 timeout 300 python bench.py --workload c2 --kind g1 --no-cpu --steps 10 > gpurun_out/s2b_bench_c2_g1_legacy.json 2>> gpurun_out/s2b_err.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s2b_bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, 'ERR', e); continue
    print(f, round(d['value']), round(d['ms_per_step'],3), 'dec', round(d.get('decode',{}).get('value',0)), d.get('decode',{}).get('pixels_match_input'), {k:round(v,3) for k,v in d.get('decode',{}).get('kernels_ms',{}).items()})
P
tail -5 gpurun_out/s2b_err.log
