# round-2 profile run: launch list + ncu --set full of the single-pass encoder (C2 RGB and C3 RGBA), bench lines of the main workloads
set -x
K='regex:encode_fused|fused_finish|fused_crc|row_scan|row_offsets|pack_rows|adler_finalize|idat_crc|huffman|row_hist|decode_|unfilter'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 80 --csv --log-file gpurun_out/launches_r2.csv python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_bench_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_fused -s 3 -c 1 -f -o gpurun_out/prof_fused_c2_r2 python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --no-decode --e2e-images 4 > gpurun_out/ncu_fused_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_fused -s 3 -c 1 -f -o gpurun_out/prof_fused_c3_r2 python bench.py --workload c3 --images 8 --steps 2 --warmup 1 --no-cpu --no-decode --e2e-images 2 > gpurun_out/ncu_fused_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_write -s 2 -c 1 -f -o gpurun_out/prof_decwrite_c2_r2 python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 --own-files > gpurun_out/ncu_decw.log 2>&1
ls -la gpurun_out/*.ncu-rep
