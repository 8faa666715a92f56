#!/bin/bash
# 8 GPUs of one box: identity checks, C2 scaling, C3 and C4 on 8, in-process F3B_DEVICES
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g8_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29601 tools/check_multigpu.py > gpurun_out/g8_check_multigpu_n8.txt 2>&1; echo "rc=$?" >> gpurun_out/g8_check_multigpu_n8.txt
tail -8 gpurun_out/g8_check_multigpu_n8.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g8_bench_c2_n1.txt 2>&1
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29610+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/g8_bench_c2_n$n.txt 2>&1
done
timeout 600 $TR --nproc-per-node 8 --master-port 29630 bench.py --gpus 8 --workload c3 --steps 3 --warmup 3 > gpurun_out/g8_bench_c3_n8.txt 2>&1
timeout 600 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/g8_bench_c3_n1.txt 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29640 bench.py --gpus 8 --workload c4 --steps 50 --warmup 5 > gpurun_out/g8_bench_c4_n8.txt 2>&1
timeout 300 python bench.py --workload c4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/g8_bench_c4_n1.txt 2>&1
F3B_DEVICES=0-7 timeout 600 python tools/measure_inproc.py > gpurun_out/g8_inproc_n8.txt 2>&1
F3B_DEVICES=0 timeout 600 python tools/measure_inproc.py c2 > gpurun_out/g8_inproc_n1.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/g8_bench_*.txt")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], "value %.4g %s  ms/step %.3f  e2e %.4g (%.3f ms)  identical %s  exact %s" % (
                d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
                d.get("image_identical_to_single_gpu"), d.get("exact_buckets")))
            break
    else:
        print(f, "NO JSON LINE:", open(f).read()[-600:])
PY
cat gpurun_out/g8_inproc_n8.txt gpurun_out/g8_inproc_n1.txt | tail -5
