#!/bin/bash
# 2 GPUs: multi-process peer path (IPC) checks + bench at N=1,2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c4_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py > gpurun_out/c4_check_multigpu_n2.txt 2>&1; echo "rc=$?" >> gpurun_out/c4_check_multigpu_n2.txt
tail -12 gpurun_out/c4_check_multigpu_n2.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench_n1.txt 2>&1; tail -c 1500 gpurun_out/c4_bench_n1.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c4_bench_n2.txt 2>&1; tail -c 2500 gpurun_out/c4_bench_n2.txt
F3B_DEVICES=0,1 timeout 300 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/c4_multi_inproc.txt 2>&1; tail -5 gpurun_out/c4_multi_inproc.txt
