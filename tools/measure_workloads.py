"""Per-stage timings of the BASELINE.json configs on one GPU (run under gpurun)."""
import os, sys, time, glob
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize, set_flame_attrs

gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()

def run(name, xml, reps=3, **kw):
    g = gpu.parse(xml)
    gpu.render(g, **kw)
    ctx.stats_reset()
    t = time.perf_counter()
    for _ in range(reps):
        img, st, _ = gpu.render(g, **kw)
    wall = (time.perf_counter() - t) / reps
    it, n1 = ctx.stage_stats(0); de, n2 = ctx.stage_stats(1); fi, n3 = ctx.stage_stats(2)
    print("%-34s wall %8.2f ms | iterate %8.3f  density %7.3f  final %7.3f ms | %.3e iters  %.3e it/s (kernel) mean %.1f"
          % (name, wall * 1e3, it / reps, de / reps, fi / reps, st["num_iters"], st["num_iters"] / (it / reps / 1e3), img.mean()), flush=True)

run("C1 test.flam3 flame0 native ts=1", set_flame_attrs(f0, temporal_samples=1))
run("C1 flame0 native ts=1000", f0)
run("C2 1080p q2000", resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1))
run("C2 1080p q200", resize(f0, 1920, 1080, scale=540, quality=200, oversample=1, temporal_samples=1))
run("C2 1080p q10 (DE heavy)", resize(f0, 1920, 1080, scale=540, quality=10, oversample=1, temporal_samples=1))
two = "<flames>" + resize(f0, 3840, 2160, scale=1080, quality=500, oversample=2, temporal_samples=1000) + \
      resize(f1, 3840, 2160, scale=1080, quality=500, oversample=2, temporal_samples=1000) + "</flames>"
run("C3 4K os2 q500 ts=1000 (1 GPU)", two, reps=2, ngenomes=2, time=50.0)
two = "<flames>" + f0 + f1 + "</flames>"
run("C4 animate frame 640x480 q10 ts=1000", two, reps=5, ngenomes=2, time=50.0)
run("C4 same, nthreads=16 (flam3-animate default)", two, reps=5, ngenomes=2, time=50.0, nthreads=16)
for p in sorted(glob.glob(os.path.join(ROOT, "tests/golden/c5_*.flam3")))[:12]:
    run("C5 " + os.path.basename(p), resize(open(p).read(), 1024, 1024, quality=1000, temporal_samples=1), reps=2)
