#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c6_c3_binned_launches.csv python tools/run_c3_once.py 500 1000 1 > gpurun_out/c6_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/c6_c3_binned_launches.csv 2>&1 | tail -20
