set -x
ncu --set full --clock-control none --import-source on -k regex:encode_fused -s 3 -c 1 -f -o gpurun_out/prof_fused_c2_r2 python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --no-decode --e2e-images 4 > gpurun_out/ncu_fused_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
