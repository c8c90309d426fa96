# pack warps-per-CTA A/B with the in-kernel CRC; combine-kernel rewrite check
set -x
timeout 600 python -m pytest tests/test_encode_gpu.py -m gpu -x -q -k "generations and two_kernel" 2>&1 | tail -3
for v in "" w5 w6 w7 w8; do
for a in "c2 g1" "c2 g0" "c3 g1" "odd g1" "c1 g0"; do set -- $a
FPNGB_LIB_VARIANT=$v timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 10 > gpurun_out/t8_$1_$2_v$v.json 2>> gpurun_out/t8_err.log
done; done
tail -3 gpurun_out/t8_err.log
