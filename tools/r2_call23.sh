#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c23_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c23_pytest_gpu.txt
grep -E "AssertionError:|^FAILED|passed|failed" gpurun_out/c23_pytest_gpu.txt | tail -15
