"""Chaos-game stage time (kernel + replica merge) against the number of histogram replicas,
at a few row pitches (a pitch change reshuffles which hot buckets share an L2 slice)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
def run(name, xml, pitches, reps=3, ks=(1, 2, 3, 4)):
    g = gpu.parse(xml)
    for p in pitches:
        row = []
        for k in ks:
            os.environ["F3B_HIST_REPLICAS"] = str(k)
            os.environ["F3B_HIST_PITCH"] = str(p)
            gpu.render(g, isaac_seed="ab"); ctx.stats_reset()
            for _ in range(reps): img, st, _ = gpu.render(g, isaac_seed="ab")
            row.append("K=%d %7.3f" % (k, ctx.stage_stats(0)[0] / reps))
        print("%-22s pitch %5d | iterate stage ms: %s" % (name, p, "  ".join(row)), flush=True)
run("C2 flame0 1080p q2000", resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), (1938, 1944, 1948, 1956, 2048, 2080))
run("flame1 1080p q2000", resize(f1, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), (1938, 1948))
run("flame0 1024^2 q1000", resize(f0, 1024, 1024, scale=256, quality=1000, oversample=1, temporal_samples=1), (1042, 1048), ks=(1, 2, 4, 6))
run("flame0 640x480 q500", resize(f0, 640, 480, scale=120, quality=500, oversample=1, temporal_samples=1), (658,), ks=(1, 2, 4, 8))
run("flame0 4K os2 q200", resize(f0, 3840, 2160, scale=1080, quality=200, oversample=2, temporal_samples=1), (7718,), reps=2, ks=(1, 2))
