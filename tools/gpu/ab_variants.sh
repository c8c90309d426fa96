# session-2 run: A/B of build variants (pack table offsets on the FMA pipe, scan kernel at 5 CTAs per SM)
for v in default ofsfma scan5 both; do
  for a in "c2 g1" "c3 g1" "c2 g0"; do set -- $a
    if [ $v = default ]; then unset FPNGB_LIB_VARIANT; else export FPNGB_LIB_VARIANT=$v; fi
    timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 20 > gpurun_out/s2f_${v}_$1_$2.json 2>> gpurun_out/s2f_err.log
  done
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s2f_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, 'ERR', e); continue
    print(f, round(d['value']), round(d['ms_per_step'],4), d['config'].get('parity_image0_vs_oracle'), {k:round(v,3) for k,v in d['kernels_ms'].items() if v})
P
tail -3 gpurun_out/s2f_err.log
