"""Chaos-game time of C2 with and without the line-permuted histogram layout, at a few row
pitches (the plain layout's time depends strongly on the pitch; the permuted one should not)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
def run(name, xml, pitches, reps=3):
    g = gpu.parse(xml)
    for p in pitches:
        row = []
        for scr in ("0", "1"):
            os.environ["F3B_HIST_SCRAMBLE"] = scr
            os.environ["F3B_HIST_PITCH"] = str(p)
            gpu.render(g, isaac_seed="ab"); ctx.stats_reset()
            for _ in range(reps): img, st, _ = gpu.render(g, isaac_seed="ab")
            it = ctx.stage_stats(0)[0] / reps; de = ctx.stage_stats(1)[0] / reps
            row.append("%s: iterate %7.3f ms density %.3f ms" % ("permuted" if scr == "1" else "plain   ", it, de))
        print("%-22s pitch %5d | %s | %s" % (name, p, row[0], row[1]), flush=True)
run("C2 flame0 1080p q2000", resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), (1938, 1944, 1948, 2048, 2080))
run("flame1 1080p q2000", resize(f1, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), (1938, 1948, 2080))
run("flame0 1024^2 q1000", resize(f0, 1024, 1024, scale=256, quality=1000, oversample=1, temporal_samples=1), (1042, 1048))
run("flame0 4K os2 q200", resize(f0, 3840, 2160, scale=1080, quality=200, oversample=2, temporal_samples=1), (7718,), reps=2)
