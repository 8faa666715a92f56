#!/bin/bash
mkdir -p gpurun_out
for b in 8 0 4 5 6 7 10 12 8; do F3B_HIST_SCRAMBLE=$b timeout 100 python tools/time_c2.py "scramble bits=$b" >> gpurun_out/c36_scramble.txt 2>&1; done
cat gpurun_out/c36_scramble.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c36_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c36_pytest_gpu.txt
grep -E "AssertionError:|^FAILED|passed|failed" gpurun_out/c36_pytest_gpu.txt | tail -8
