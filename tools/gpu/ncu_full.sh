set -x
for k in idat_crc_stream decode_write decode_scan; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_r2 python bench.py --images 64 --steps 2 --warmup 1 --no-cpu --e2e-images 4 --own-files > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*_r2.ncu-rep
