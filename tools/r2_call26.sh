#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c26_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c26_pytest_gpu.txt
grep -E "AssertionError:|^FAILED|passed|failed" gpurun_out/c26_pytest_gpu.txt | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/r02_bench_launches.csv | tee gpurun_out/r02_bench_launch_shares.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:f3b_iterate_kernel -s 3 -c 1 -f -o gpurun_out/r02_iterate python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_iterate_ncu.log 2>&1
tail -2 gpurun_out/r02_iterate_ncu.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
