#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_sweep.py > gpurun_out/c18_ab_sweep.txt 2>&1; head -9 gpurun_out/c18_ab_sweep.txt
timeout 900 python tools/ab_sweep_big.py > gpurun_out/c18_ab_sweep_big.txt 2>&1; tail -14 gpurun_out/c18_ab_sweep_big.txt
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/c18_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c18_pytest_gpu.txt
tail -8 gpurun_out/c18_pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c18_bench_n1.txt 2>&1; tail -c 300 gpurun_out/c18_bench_n1.txt; grep -o '"ms_per_step": [0-9.]*' gpurun_out/c18_bench_n1.txt
