"""Does the bucket-row pitch matter (L2 slice balance)? C2 at slightly different widths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
for w in (1920, 1902, 1910, 1912, 1918, 1926, 1934, 1950, 1982, 2030, 2046, 2048):
    g = gpu.parse(resize(f0, w, 1080, scale=540, quality=2000 * 1920.0 / w, oversample=1, temporal_samples=1))
    gpu.render(g); ctx.stats_reset()
    for _ in range(3): _, st, _ = gpu.render(g)
    ms = ctx.stage_stats(0)[0] / 3
    print("width %4d (pitch %5d buckets, %6d B): iterate %.3f ms  %.3e it/s" % (w, w + 18, (w + 18) * 16, ms, st["num_iters"] / ms * 1e3), flush=True)
