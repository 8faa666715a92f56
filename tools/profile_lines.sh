#!/bin/bash
# Run under gpurun: per-instruction execution counts of one chaos-game launch.
mkdir -p gpurun_out
ncu --section SourceCounters --section InstructionStats --import-source on --clock-control none -k regex:f3b_iterate -s 3 -c 1 \
    -o gpurun_out/iter_src -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_src.log 2>&1
tail -2 gpurun_out/ncu_src.log | cut -c1-200
