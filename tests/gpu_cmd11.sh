set -x
for a in "c2 g1" "c2 g0" "c3 g1"; do set -- $a
timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --no-decode --steps 10 > gpurun_out/t11_$1_$2.json 2>> gpurun_out/t11_err.log
done
for k in decode_scan decode_write pack_rows16; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${k}_c2_r2b python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*r2b.ncu-rep
