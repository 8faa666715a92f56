#!/bin/bash
mkdir -p gpurun_out
cp flam3_b200/lib/libflam3_b200.so /tmp/keep.so
for rnd in 1; do
for v in pickidx final scr20 scr12 scr8; do
  cp variants/lib_$v.so flam3_b200/lib/libflam3_b200.so
  timeout 300 python tools/time_c2.py "$v" >> gpurun_out/c35_variants.txt 2>&1
done; done
cp /tmp/keep.so flam3_b200/lib/libflam3_b200.so
cat gpurun_out/c35_variants.txt
