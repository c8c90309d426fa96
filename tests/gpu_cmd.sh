set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_g1.json 2> gpurun_out/bench_c2_g1.err; tail -3 gpurun_out/bench_c2_g1.err; cat gpurun_out/bench_c2_g1.json
python bench.py --steps 10 --warmup 3 --kind g0 --no-cpu > gpurun_out/bench_c2_g0.json 2>> gpurun_out/bench_c2_g1.err; cat gpurun_out/bench_c2_g0.json
K='regex:row_scan|row_offsets|pack_rows|adler_finalize|idat_crc|huffman'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/launches_r1.csv python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_bench.log 2>&1
for k in row_scan pack_rows idat_crc; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${k}_r1 python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
