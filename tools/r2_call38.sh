#!/bin/bash
mkdir -p gpurun_out
cp flam3_b200/lib/libflam3_b200.so /tmp/keep.so
for rnd in 1 2; do
for v in float4 rgba8; do
  cp variants/lib_$v.so flam3_b200/lib/libflam3_b200.so
  timeout 300 python tools/time_genomes.py "$v" >> gpurun_out/c38_variants.txt 2>&1
done; done
cp /tmp/keep.so flam3_b200/lib/libflam3_b200.so
F3B_WU=1 timeout 300 python tools/time_genomes.py "float4 F3B_WU=1" >> gpurun_out/c38_variants.txt 2>&1
cat gpurun_out/c38_variants.txt
