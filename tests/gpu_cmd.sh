set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --workload c4 --no-cpu --steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','kernels_ms','decode')}))"
K='regex:row_scan|row_offsets|pack_rows|adler_finalize|idat_crc|huffman|row_hist|decode_|unfilter'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --e2e-images 4 > gpurun_out/ncu_bench.log 2>&1
for k in row_scan16 pack_rows16 idat_crc; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${k}_r1_final python bench.py --images 32 --steps 2 --warmup 1 --no-cpu --no-decode --e2e-images 4 > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:decode_write -s 2 -c 1 -f -o gpurun_out/prof_decode_write_r1_final python tests/ncu_decode_driver.py c2 32 > gpurun_out/ncu_decw.log 2>&1
ls gpurun_out | wc -l
