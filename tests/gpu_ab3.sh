set -x
timeout 600 python -m pytest tests/test_decode_gpu.py -m gpu -x -q -k "write_paths or corrupted or batch" --durations=5 2>&1 | tail -12 > gpurun_out/s2e_pytest.txt; cat gpurun_out/s2e_pytest.txt
for a in "c2 g1" "c3 g1" "c4 g1"; do set -- $a; timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --steps 10 > gpurun_out/s2e_bench_$1_$2.json 2>> gpurun_out/s2e_err.log; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s2e_bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, 'ERR', e); continue
    print(f, round(d['value']), round(d['ms_per_step'],3), 'dec', round(d.get('decode',{}).get('value',0)), d.get('decode',{}).get('ms_per_step'), d.get('decode',{}).get('pixels_match_input'), {k:round(v,3) for k,v in d.get('decode',{}).get('kernels_ms',{}).items()})
P
tail -5 gpurun_out/s2e_err.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_write_staged -s 1 -c 1 -f -o gpurun_out/prof_ws_c2_v2 python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_ws2.log 2>&1
