timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_reference_fixtures_gpu.py::test_example_png_decode_gpu tests/test_encode_gpu.py::test_full_size_configs_properties -x -q 2>&1 | tail -5
for a in "c2 g1" "c3 g1" "c2 g0"; do set -- $a
timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1 $2 enc', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d.get('kernels_ms').items() if v}, d['config'].get('parity_image0_vs_oracle')); print('   decode', round(d['decode']['value']), round(d['decode']['ms_per_step'],3), {k: round(v,3) for k,v in d['decode'].get('kernels_ms').items()}, d['decode']['pixels_match_input'], 'e2e', round(d['e2e']['value']), 'dec e2e', round(d['decode']['e2e']['value']))"
done
