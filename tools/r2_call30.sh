#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_c2.py "trimmed" > gpurun_out/c30_time_c2.txt 2>&1; cat gpurun_out/c30_time_c2.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c30_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c30_pytest_gpu.txt
grep -E "AssertionError:|^FAILED|passed|failed" gpurun_out/c30_pytest_gpu.txt | tail -8
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/c30_sanitize_racecheck.txt 2>&1
grep -E "RACECHECK SUMMARY|done|Traceback" gpurun_out/c30_sanitize_racecheck.txt | tail -3
