#!/bin/bash
# Run under gpurun. Produces gpurun_out/launches.csv (every launch with its device time)
# and gpurun_out/iter_full.ncu-rep (--set full capture of the chaos-game kernel).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:f3b_iterate -s 3 -c 1 \
    -o gpurun_out/iter_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
