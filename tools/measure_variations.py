"""Chaos-game throughput per variation (config C5 of SURVEY.md §8(d), performance side):
the single-variation genomes of tests/variation_cases.py at 1024x1024 quality 1000 on one
GPU. Run under gpurun; writes gpurun_out/variations_perf.txt."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import flam3_b200 as fb
from variation_cases import VAR_NAMES, genome_xml

gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
rows = []
for vid in range(99):
    g = gpu.parse(genome_xml(vid, size=1024, quality=1000))
    gpu.render(g, isaac_seed="perf")
    ctx.stats_reset()
    reps = 2
    for _ in range(reps):
        img, st, _ = gpu.render(g, isaac_seed="perf")
    it, _ = ctx.stage_stats(0)
    rows.append((vid, VAR_NAMES[vid], st["num_iters"] / (it / reps / 1e3), it / reps, st["badvals"], img.mean()))
rows.sort(key=lambda r: r[2])
with open(os.path.join(ROOT, "gpurun_out", "variations_perf.txt"), "w") as f:
    f.write("# vid name iters_per_s iterate_ms badvals image_mean (1024x1024 q1000, 1.05e9 iterations)\n")
    for r in rows:
        line = "%2d %-14s %.3e %8.3f %10d %6.1f" % r
        f.write(line + "\n")
        print(line)
