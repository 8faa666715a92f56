#!/bin/bash
mkdir -p gpurun_out
cp flam3_b200/lib/libflam3_b200.so /tmp/keep.so
for rnd in 1 2; do
for v in norep report; do
  cp variants/lib_$v.so flam3_b200/lib/libflam3_b200.so
  timeout 300 python tools/time_c2.py "$v" >> gpurun_out/c19_variants.txt 2>&1
done; done
cp /tmp/keep.so flam3_b200/lib/libflam3_b200.so
cat gpurun_out/c19_variants.txt
F3B_TIMING=1 timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/c19_c4.txt 2> gpurun_out/c19_c4_timing.txt; tail -6 gpurun_out/c19_c4_timing.txt
