#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exact.py -q -x > gpurun_out/c5_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.txt
tail -30 gpurun_out/c5_pytest.txt
timeout 600 python tools/measure_c3.py > gpurun_out/c5_c3.txt 2>&1; tail -20 gpurun_out/c5_c3.txt
