#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/c10_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.txt
tail -25 gpurun_out/c10_pytest.txt
bash tools/r2_profile.sh
