"""Chaos-game stage time of a few 1080p q2000 frames (flame 0 and 1 of test.flam3, generic
kernel, a C5 genome), for same-box A/B of layout knobs given in the environment."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize, rescale
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
rd = lambda p: open(os.path.join(ROOT, "tests/golden", p)).read()
cases = [("flame0", resize(rd("test_flame0.flam3"), 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), {}),
         ("flame1", resize(rd("test_flame1.flam3"), 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), {}),
         ("flame0 generic", resize(rd("test_flame0.flam3"), 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), {"F3B_NO_SIMPLE": "1"}),
         ("c5 var00 1024 q2000", rescale(rd("c5/c5_var00.flam3"), 1024, 1024, quality=2000), {}),
         ("c5 mix110 1024 q1000", rescale(rd("c5/c5_mix110.flam3"), 1024, 1024, quality=1000), {})]
out = []
for name, xml, env in cases:
    os.environ.pop("F3B_NO_SIMPLE", None)
    os.environ.update(env)
    g = gpu.parse(xml)
    gpu.render(g, isaac_seed="ab"); ctx.stats_reset()
    for _ in range(5): gpu.render(g, isaac_seed="ab")
    out.append("%s %.3f" % (name, ctx.stage_stats(0)[0] / 5))
print((sys.argv[1] if len(sys.argv) > 1 else "") + " | " + " | ".join(out), flush=True)
