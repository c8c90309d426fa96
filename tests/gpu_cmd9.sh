# decoder 128-bit stores + pack pixel skip: full GPU suite + decode numbers
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for a in "c2 g1" "c2 g0" "c3 g1" "c4 g1"; do set -- $a
timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --steps 10 > gpurun_out/t13_$1_$2.json 2>> gpurun_out/t13_err.log
done
tail -3 gpurun_out/t13_err.log
