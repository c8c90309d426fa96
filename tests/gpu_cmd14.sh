# multiplicative bit stager A/B
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "" mul0; do
for a in "c2 g1" "c2 g0" "c3 g1" "c4 g1" "odd g1" "c2 g2"; do set -- $a
FPNGB_LIB_VARIANT=$v timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 10 > gpurun_out/t14_$1_$2_v$v.json 2>> gpurun_out/t14_err.log
done; done
tail -3 gpurun_out/t14_err.log
