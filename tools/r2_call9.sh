#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_analysis.py tests/test_gpu_postpass.py tests/test_gpu_regressions.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/c9_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.txt
tail -30 gpurun_out/c9_pytest.txt
timeout 600 python tools/ab_post.py > gpurun_out/c9_ab_post.txt 2>&1; tail -20 gpurun_out/c9_ab_post.txt
