#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_c5.py -m gpu -q -x -k "warp_uniform" > gpurun_out/c12_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c12_pytest.txt
tail -15 gpurun_out/c12_pytest.txt
timeout 900 python tools/measure_c5.py > gpurun_out/c12_c5_perf.txt 2>&1; tail -8 gpurun_out/c12_c5_perf.txt
timeout 300 python tools/ab_post.py > gpurun_out/c12_ab_post.txt 2>&1; grep -E "new" gpurun_out/c12_ab_post.txt
