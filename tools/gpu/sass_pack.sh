# dev helper: compile encode16_kernels.cu to a cubin and dump the SASS of pack/scan kernels into gpurun_out/sass/
set -e
cd /root/repo/fpng_b200/csrc
mkdir -p /root/repo/gpurun_out/sass
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin -o /root/repo/gpurun_out/sass/enc16.cubin encode16_kernels.cu -Xptxas -v 2>&1 | grep -E "error|registers" || true
for k in pack_rows16_kernelILi3 pack_rows16_kernelILi4 row_scan16_kernelILi3 row_scan16_kernelILi4; do
  f=$(cuobjdump -elf /root/repo/gpurun_out/sass/enc16.cubin | grep -o "_ZN5fpngb[0-9]*${k}[A-Za-z0-9_]*" | head -1)
  cuobjdump -sass -fun "$f" /root/repo/gpurun_out/sass/enc16.cubin | grep -v "^\s*/\* 0x" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s#^\s+/\*([0-9a-f]+)\*/\s+#\1 #' | cut -c1-90 > /root/repo/gpurun_out/sass/$k.txt
  echo $k $(wc -l < /root/repo/gpurun_out/sass/$k.txt)
done
