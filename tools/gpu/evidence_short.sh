# round-2 third-session evidence run of HEAD (1 GPU, <= 9 GPU-minutes left): ordered by priority, every step under its own timeout
set -x
mkdir -p gpurun_out
T0=$(date +%s)
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.txt 2>&1; tail -1 gpurun_out/r2g_smoke.txt
timeout 200 python bench.py > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_err.log; tail -c 600 gpurun_out/r2g_bench_c2.json
echo "t=$(( $(date +%s) - T0 ))"
B="python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4"
timeout 170 ncu --set full --clock-control none --import-source on -k "regex:row_scan16_kernel|pack_rows16_kernel" -s 6 -c 2 -f -o gpurun_out/prof_encode_c2_r2g $B > gpurun_out/ncu_encode.log 2>&1
echo "t=$(( $(date +%s) - T0 ))"
timeout 170 ncu --set full --clock-control none --import-source on -k "regex:decode_scan_kernel|decode_write_kernel" -s 6 -c 2 -f -o gpurun_out/prof_decode_c2_r2g $B > gpurun_out/ncu_decode.log 2>&1
echo "t=$(( $(date +%s) - T0 ))"
K='regex:row_scan|row_offsets|pack_rows|adler_finalize|idat_crc|row_crc|huffman|row_hist|decode_|unfilter'
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 80 --csv --log-file gpurun_out/r2g_launches_c2_32img.csv $B > gpurun_out/ncu_bench_r2g.log 2>&1
echo "t=$(( $(date +%s) - T0 ))"
ls -la gpurun_out/*r2g.ncu-rep
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_reference_arm.json 2>> gpurun_out/r2g_bench_err.log
timeout 120 python bench.py --workload c3 --kind g1 --no-cpu --steps 10 > gpurun_out/r2g_bench_c3_g1.json 2>> gpurun_out/r2g_bench_err.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2g_pytest_gpu.txt; cat gpurun_out/r2g_pytest_gpu.txt
echo "t=$(( $(date +%s) - T0 ))"
for a in "odd g1" "c4 g1" "c2 g0"; do set -- $a; timeout 100 python bench.py --workload $1 --kind $2 --no-cpu --steps 10 > gpurun_out/r2g_bench_$1_$2.json 2>> gpurun_out/r2g_bench_err.log; done
echo "t=$(( $(date +%s) - T0 ))"
