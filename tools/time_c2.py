"""Chaos-game stage time of C2 (1080p q2000), sweepers on and off: for same-box A/B of two builds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
g = gpu.parse(resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1))
out = []
for name, env in (("sweep off", {"F3B_SWEEP": "0"}), ("default", {}), ("sweep off", {"F3B_SWEEP": "0"}), ("default", {})):
    os.environ.pop("F3B_SWEEP", None)
    os.environ.update(env)
    gpu.render(g, isaac_seed="ab"); ctx.stats_reset()
    for _ in range(8): gpu.render(g, isaac_seed="ab")
    out.append("%s %.3f ms" % (name, ctx.stage_stats(0)[0] / 8))
print((sys.argv[1] if len(sys.argv) > 1 else "") + "  " + "   ".join(out), flush=True)
