"""One process, several GPUs (F3B_DEVICES): C2 / C3-shaped frame times through flam3_render,
host output buffer included. Run as: F3B_DEVICES=0-7 python tools/measure_inproc.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
def c3(q, ts):
    return "<flames>" + resize(f0, 3840, 2160, scale=1080, quality=q, oversample=2, temporal_samples=ts) + \
      resize(f1, 3840, 2160, scale=1080, quality=q, oversample=2, temporal_samples=ts).replace('time="0"', 'time="100"', 1) + "</flames>"
cases = [("C2 1080p q2000", resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), {}, 10),
         ("C3 4K os2 q5000 ts=1000", c3(5000, 1000), dict(ngenomes=2, time=50.0), 2)]
if len(sys.argv) > 1 and sys.argv[1] == "c2":
    cases = cases[:1]
for name, xml, kw, reps in cases:
    g = gpu.parse(xml)
    img, st, _ = gpu.render(g, isaac_seed="inproc", **kw)
    img, st, _ = gpu.render(g, isaac_seed="inproc", **kw)
    t = time.perf_counter()
    for _ in range(reps):
        img, st, _ = gpu.render(g, isaac_seed="inproc", **kw)
    dt = (time.perf_counter() - t) / reps
    ctxs = []
    i = 0
    while True:
        c = gpu.render_context_at(core, i)
        if c is None:
            break
        ctxs.append(c)
        i += 1
    print("%-26s F3B_DEVICES=%s (%d contexts): %.2f ms per flam3_render call, %.3e iters/s, mean %.2f, lost %s" % (
        name, os.environ.get("F3B_DEVICES", "-"), len(ctxs), dt * 1e3, st["num_iters"] / dt, img.mean(),
        [c.exactness()[0] for c in ctxs]), flush=True)
