#!/bin/bash
mkdir -p gpurun_out
for m in 1 0 1 0; do
  if [ $m = 1 ]; then export F3B_INTERP_REF=1; else unset F3B_INTERP_REF; fi
  echo "F3B_INTERP_REF=$m $(timeout 200 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1)" >> gpurun_out/c39_c4.txt
done
unset F3B_INTERP_REF
cat gpurun_out/c39_c4.txt
timeout 600 python -m pytest tests/test_gpu_cli.py tests/test_gpu_frontend.py tests/test_gpu_chaos.py -m gpu -q 2>&1 | tail -2
