#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_c2.py "single" > gpurun_out/c27_time_c2.txt 2>&1; cat gpurun_out/c27_time_c2.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c27_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c27_pytest_gpu.txt
grep -E "AssertionError:|^FAILED|passed|failed" gpurun_out/c27_pytest_gpu.txt | tail -8
timeout 600 python tools/measure_c5.py > gpurun_out/c27_c5.txt 2>&1; tail -5 gpurun_out/c27_c5.txt
