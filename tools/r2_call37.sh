#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 3" "0 3" "8 2" "8 4" "7 3"; do set -- $cfg; F3B_HIST_SCRAMBLE=$1 F3B_HIST_REPLICAS=$2 timeout 200 python tools/time_genomes.py "scramble=$1 replicas=$2" >> gpurun_out/c37_genomes.txt 2>&1; done
cat gpurun_out/c37_genomes.txt
