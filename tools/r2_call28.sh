#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/c28_bench_c2.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/c28_bench_c2.txt
timeout 600 python bench.py --impl reference > gpurun_out/c28_bench_ref.txt 2>&1; grep -o '"value": [0-9.e+]*' gpurun_out/c28_bench_ref.txt | head -1
timeout 600 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c28_bench_c3.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/c28_bench_c3.txt
timeout 300 python bench.py --workload c4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/c28_bench_c4.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/c28_bench_c4.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c28_smoke.txt 2>&1; tail -2 gpurun_out/c28_smoke.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:f3b_iterate_kernel -s 3 -c 1 -f -o gpurun_out/r02_iterate python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_iterate_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/r02_bench_launches.csv | tee gpurun_out/r02_bench_launch_shares.txt | head -4
