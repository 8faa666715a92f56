"""A/B on one box: post-pass stage times with the closed-form DE gather / float-first final
pass against the table gather / all-double final pass (F3B_DE_TABLE, F3B_FINAL_DOUBLE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
cases = [("C2 1080p q2000", resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), {}),
         ("1080p q200", resize(f0, 1920, 1080, scale=540, quality=200, oversample=1, temporal_samples=1), {}),
         ("1080p q10", resize(f0, 1920, 1080, scale=540, quality=10, oversample=1, temporal_samples=1), {}),
         ("C1 640x480 q10", resize(f0, 640, 480, scale=240, quality=10, temporal_samples=1), {}),
         ("4K os2 still q100", resize(f0, 3840, 2160, scale=1080, quality=100, oversample=2, temporal_samples=1), {}),
         ("C3 q200 ts=1000", "<flames>" + resize(f0, 3840, 2160, scale=1080, quality=200, oversample=2, temporal_samples=1000) +
          resize(f1, 3840, 2160, scale=1080, quality=200, oversample=2, temporal_samples=1000).replace('time="0"', 'time="100"', 1) + "</flames>",
          dict(ngenomes=2, time=50.0))]
for name, xml, kw in cases:
    g = gpu.parse(xml)
    imgs = {}
    for tag, env in (("new", {}), ("table+double", {"F3B_DE_TABLE": "1", "F3B_FINAL_DOUBLE": "1"})):
        for k in ("F3B_DE_TABLE", "F3B_FINAL_DOUBLE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        gpu.render(g, isaac_seed="ab", **kw); ctx.stats_reset()
        for _ in range(3):
            img, st, _ = gpu.render(g, isaac_seed="ab", **kw)
        de, nde = ctx.stage_stats(1); fi, nfi = ctx.stage_stats(2)
        imgs[tag] = img
        print("%-20s %-13s density %7.3f ms  final %7.3f ms" % (name, tag, de / 3, fi / 3), flush=True)
    d = np.abs(imgs["new"].astype(int) - imgs["table+double"].astype(int))
    print("%-20s bytes differing between the two: %d of %d (max %d)" % (name, int((d > 0).sum()), d.size, int(d.max())), flush=True)
