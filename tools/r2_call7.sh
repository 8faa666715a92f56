#!/bin/bash
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:f3b_iterate -c 1 -o gpurun_out/c7_iter_binned -f python tools/run_c3_once.py 200 1000 1 > gpurun_out/c7_ncu.log 2>&1
tail -3 gpurun_out/c7_ncu.log
ls -la gpurun_out/c7_iter_binned.ncu-rep
