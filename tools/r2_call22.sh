#!/bin/bash
mkdir -p gpurun_out
F3B_TIMING=1 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/c22_c4_timing.txt | grep -o '"ms_per_step": [0-9.]*' > gpurun_out/c22_c4.txt
grep "host ms): setup" gpurun_out/c22_c4_timing.txt | tail -6 >> gpurun_out/c22_c4.txt; cat gpurun_out/c22_c4.txt
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/c22_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c22_pytest_gpu.txt
tail -5 gpurun_out/c22_pytest_gpu.txt
