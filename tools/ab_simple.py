"""A/B on one box: C2 chaos-game time with the SIMPLE kernel and with the generic one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
g = gpu.parse(resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1))
def t(reps=5):
    gpu.render(g); ctx.stats_reset()
    for _ in range(reps): gpu.render(g)
    return ctx.stage_stats(0)[0] / reps
for rnd in range(3):
    os.environ.pop("F3B_NO_SIMPLE", None); a = t()
    os.environ["F3B_NO_SIMPLE"] = "1"; b = t()
    print("round %d: simple %.3f ms   generic %.3f ms" % (rnd, a, b), flush=True)
