#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c15_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c15_pytest.txt
tail -12 gpurun_out/c15_pytest.txt
timeout 900 python tools/measure_c5.py > gpurun_out/c15_c5_perf.txt 2>&1; tail -4 gpurun_out/c15_c5_perf.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c15_bench.txt 2>&1; python - <<'PY'
import json
for l in open("gpurun_out/c15_bench.txt"):
    if l.startswith("{"):
        d = json.loads(l); print("bench C2: %.3f ms/step, e2e %.3f ms, kernel %.3f, post %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["avg_launch_ms"], [round(p["stage_ms"], 3) for p in d["roofline_post"]]))
PY
