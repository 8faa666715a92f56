#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/psnr_spread.py 27 14 45 > gpurun_out/c21_psnr_spread.txt 2>&1; cat gpurun_out/c21_psnr_spread.txt | tail -8
for m in 1 0 1 0; do
  if [ $m = 1 ]; then export F3B_INTERP_REF=1; else unset F3B_INTERP_REF; fi
  echo "F3B_INTERP_REF=$m" >> gpurun_out/c21_c4.txt
  F3B_TIMING=1 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/c21_c4_timing_$m.txt | grep -o '"ms_per_step": [0-9.]*' >> gpurun_out/c21_c4.txt
  grep "host ms): setup" gpurun_out/c21_c4_timing_$m.txt | tail -3 >> gpurun_out/c21_c4.txt
done
unset F3B_INTERP_REF
cat gpurun_out/c21_c4.txt
timeout 300 python tools/time_c2.py "mode0" > gpurun_out/c21_time_c2.txt 2>&1; cat gpurun_out/c21_time_c2.txt
