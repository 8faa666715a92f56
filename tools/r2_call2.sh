#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.txt
tail -30 gpurun_out/c2_pytest.txt
timeout 300 python tools/ab_sweep.py > gpurun_out/c2_ab_sweep.txt 2>&1; tail -20 gpurun_out/c2_ab_sweep.txt
