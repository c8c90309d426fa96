N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), d['config']['workload'])
print('gather', d.get('gather'))
print('value_with_gather', d.get('value_with_gather'))
print('e2e', d.get('e2e'))
print('decode', {k: d['decode'][k] for k in ('value','ms_per_step','input','pixels_match_input')} if 'decode' in d else None)
PY
