#!/bin/bash
# round 2 profiles: launch list of the bench command, full captures of the chaos-game kernel
# (C2) and of the post-pass kernels at C3 size
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/r02_bench_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:f3b_iterate_kernel -s 3 -c 1 -f -o gpurun_out/r02_iterate python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_iterate_ncu.log 2>&1
tail -2 gpurun_out/r02_iterate_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'f3b_de_|f3b_final|f3b_hist_merge|f3b_hist_reduce' -s 8 -c 8 -f -o gpurun_out/r02_post_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_post_c2_ncu.log 2>&1
tail -2 gpurun_out/r02_post_c2_ncu.log
F3B_BINNED=0 timeout 900 ncu --set full --clock-control none -k regex:'f3b_de_|f3b_final' -s 3 -c 4 -f -o gpurun_out/r02_post_c3 python tools/run_c3_once.py 200 1000 2 > gpurun_out/r02_post_c3_ncu.log 2>&1
tail -2 gpurun_out/r02_post_c3_ncu.log
ls -la gpurun_out/*.ncu-rep
