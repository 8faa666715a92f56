"""A/B on one box: C2 chaos-game time with and without the warp-aggregation check."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
for name, f in (("flame0", f0), ("flame1", f1)):
    g = gpu.parse(resize(f, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1))
    def t(reps=4):
        gpu.render(g); ctx.stats_reset()
        for _ in range(reps): gpu.render(g)
        return ctx.stage_stats(0)[0] / reps
    for rnd in range(2):
        os.environ.pop("F3B_AGG", None); a = t()
        os.environ["F3B_AGG"] = "0"; b = t()
        print("%s round %d: probed %.3f ms   forced off %.3f ms" % (name, rnd, a, b), flush=True)
    os.environ.pop("F3B_AGG", None)
