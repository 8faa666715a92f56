#!/bin/bash
# full GPU suite + the default bench (both arms) + smoke, as the driver runs them at round end
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/c16_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/c16_pytest_gpu.txt
tail -25 gpurun_out/c16_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c16_smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/c16_smoke.txt; tail -3 gpurun_out/c16_smoke.txt
timeout 600 python bench.py > gpurun_out/c16_bench_n1.txt 2>&1; tail -c 2500 gpurun_out/c16_bench_n1.txt
timeout 600 python bench.py --impl reference > gpurun_out/c16_bench_ref.txt 2>&1; tail -c 1200 gpurun_out/c16_bench_ref.txt
