#!/bin/bash
mkdir -p gpurun_out
cp flam3_b200/lib/libflam3_b200.so /tmp/keep.so
for v in pickidx final; do
  cp variants/lib_$v.so flam3_b200/lib/libflam3_b200.so
  timeout 600 ncu --set full --clock-control none -k regex:f3b_iterate_kernel -s 3 -c 1 -f -o gpurun_out/c34_$v python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c34_$v.log 2>&1
  for r in 3 4; do F3B_HIST_REPLICAS=$r timeout 100 python tools/time_c2.py "$v replicas=$r" >> gpurun_out/c34_replicas.txt 2>&1; done
done
cp /tmp/keep.so flam3_b200/lib/libflam3_b200.so
cat gpurun_out/c34_replicas.txt; ls -la gpurun_out/c34_*.ncu-rep
