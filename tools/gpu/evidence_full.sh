# round-2 final evidence run (1 GPU): full GPU suite, bench lines of every workload, reference arm, launch list, ncu of the dominant kernels, sanitizers
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2g_pytest_gpu.txt; cat gpurun_out/r2g_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_err.log; tail -2 gpurun_out/r2g_bench_err.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_reference_arm.json 2>> gpurun_out/r2g_bench_err.log
for a in "c3 g1" "c2 g0" "c2 g2" "c4 g1" "c1 g0" "odd g1"; do set -- $a; timeout 300 python bench.py --workload $1 --kind $2 --no-cpu --steps 10 > gpurun_out/r2g_bench_$1_$2.json 2>> gpurun_out/r2g_bench_err.log; done
K='regex:row_scan|row_offsets|pack_rows|adler_finalize|idat_crc|row_crc|huffman|row_hist|decode_|unfilter'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 80 --csv --log-file gpurun_out/r2g_launches_c2_32img.csv python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_bench_r2g.log 2>&1
for k in row_scan16 pack_rows16 decode_scan decode_write; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 3 -c 1 -f -o gpurun_out/prof_${k}_c2_r2g python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*r2g.ncu-rep
for t in memcheck racecheck; do
  timeout 200 compute-sanitizer --tool $t --error-exitcode 3 python tests/sanitize_driver.py > gpurun_out/r2g_sanitize_$t.log 2>&1; echo "sanitize $t rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|sanitize_driver" gpurun_out/r2g_sanitize_$t.log | tail -3
done
