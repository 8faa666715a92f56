#!/bin/bash
# round 2, GPU call 1: parity suite, sweeper A/B, workload table, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.txt
tail -15 gpurun_out/c1_pytest.txt
timeout 300 python tools/ab_sweep.py > gpurun_out/c1_ab_sweep.txt 2>&1; tail -20 gpurun_out/c1_ab_sweep.txt
timeout 400 python tools/measure_workloads.py > gpurun_out/c1_workloads.txt 2>&1; tail -25 gpurun_out/c1_workloads.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench.txt 2>&1; tail -3 gpurun_out/c1_bench.txt
