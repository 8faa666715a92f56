#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.txt
tail -30 gpurun_out/c3_pytest.txt
timeout 300 python tools/ab_sweep.py > gpurun_out/c3_ab_sweep.txt 2>&1; tail -24 gpurun_out/c3_ab_sweep.txt
