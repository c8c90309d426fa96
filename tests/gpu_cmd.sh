set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_r1_c2_final.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
for a in "c3 g1" "c2 g0" "c3 g0" "c2 g2" "c4 g1" "c1 g0"; do set -- $a; timeout 200 python bench.py --workload $1 --kind $2 --no-cpu --steps 10 > gpurun_out/bench_r1_$1_$2_final.json 2>> gpurun_out/bench_err.log; done
K='regex:row_scan|row_offsets|pack_rows|adler_finalize|idat_crc|huffman|row_hist|decode_|unfilter'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_bench.log 2>&1
for k in row_scan16 pack_rows16; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${k}_r1_final python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --no-decode --e2e-images 4 > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:pack_rows16 -s 3 -c 1 -f -o gpurun_out/prof_pack_rows16_c3_r1_final python bench.py --workload c3 --images 8 --steps 2 --warmup 1 --no-cpu --no-decode --e2e-images 2 > gpurun_out/ncu_pack_c3.log 2>&1
tail -2 gpurun_out/bench_err.log
for t in memcheck racecheck synccheck; do
  timeout 240 compute-sanitizer --tool $t --error-exitcode 3 python tests/sanitize_driver.py > gpurun_out/sanitize_$t.log 2>&1; echo "sanitize $t rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$t.log | tail -2
done
