#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_palette.py > gpurun_out/c17_ab_palette.txt 2>&1; cat gpurun_out/c17_ab_palette.txt | tail -12
timeout 900 python tools/ab_sweep_big.py > gpurun_out/c17_ab_sweep_big.txt 2>&1; tail -14 gpurun_out/c17_ab_sweep_big.txt
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/c17_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c17_pytest_gpu.txt
tail -8 gpurun_out/c17_pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c17_bench_n1.txt 2>&1; tail -c 600 gpurun_out/c17_bench_n1.txt
