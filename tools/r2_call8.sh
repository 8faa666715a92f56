#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 > gpurun_out/c8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.txt
tail -40 gpurun_out/c8_pytest.txt
timeout 600 python tools/ab_post.py > gpurun_out/c8_ab_post.txt 2>&1; tail -30 gpurun_out/c8_ab_post.txt
