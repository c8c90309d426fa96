# pack-kernel CRC: full GPU suite, then A/B bench lines
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for e in default two_kernel_file_crc; do
timeout 200 python bench.py --no-cpu --no-decode --steps 10 --encoder $e > gpurun_out/t7_c2_g1_$e.json 2>> gpurun_out/t7_err.log
done
for a in "c3 g1" "c2 g0" "c2 g2" "c4 g1" "c1 g0" "odd g1"; do set -- $a; timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 10 > gpurun_out/t7_$1_$2.json 2>> gpurun_out/t7_err.log; done
tail -3 gpurun_out/t7_err.log
