#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_c2.py "mode0" > gpurun_out/c20_time_c2.txt 2>&1; cat gpurun_out/c20_time_c2.txt
F3B_TIMING=1 timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/c20_c4.txt 2> gpurun_out/c20_c4_timing.txt; tail -3 gpurun_out/c20_c4_timing.txt
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/c20_c4_bench.txt 2>&1; grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/c20_c4_bench.txt | head -4
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/c20_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c20_pytest_gpu.txt
tail -5 gpurun_out/c20_pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c20_bench_n1.txt 2>&1; grep -o '"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/c20_bench_n1.txt
