"""A/B on one box: the warp-uniform xform step (F3B_WU=1, default) against per-lane picks
(F3B_WU=0): chaos-game time on C2 and on heavy random genomes (C5), and the Monte-Carlo noise
of the result against the unmodified reference (same bound as the parity tests)."""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
from parity_util import psnr
gpu, core = fb.load_dropin(), fb.load_core()
ref = fb.Flam3Lib(os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so"))
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
def timed(g, reps=3, **kw):
    gpu.render(g, **kw); ctx.stats_reset()
    for _ in range(reps): img, st, _ = gpu.render(g, **kw)
    return ctx.stage_stats(0)[0] / reps, img, st
cases = [("C2 1080p q2000", resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1))]
for name in ("mix107", "mix116", "mix123", "mix101", "mix110", "var33", "var13", "var50", "var76", "var20", "var45"):
    p = sorted(glob.glob(os.path.join(ROOT, "tests/golden/c5", "*%s.flam3" % name)))
    if p: cases.append(("C5 " + name, resize(open(p[0]).read(), 1024, 1024, quality=1000, temporal_samples=1)))
for name, xml in cases:
    g = gpu.parse(xml)
    out = {}
    for wu in ("0", "1"):
        os.environ["F3B_WU"] = wu
        ms, img, st = timed(g, isaac_seed="wu")
        out[wu] = (ms, img, st)
    print("%-16s per-lane %8.3f ms (%.3e it/s)   warp-uniform %8.3f ms (%.3e it/s)   x%.2f" % (
        name, out["0"][0], out["0"][2]["num_iters"] / out["0"][0] * 1e3, out["1"][0],
        out["1"][2]["num_iters"] / out["1"][0] * 1e3, out["0"][0] / out["1"][0]), flush=True)
# noise against the reference at equal quality (small frames, like the parity tests)
for name, xml in [("flame0 320x240 q50", resize(f0, 320, 240, scale=120, quality=50, temporal_samples=1)),
                  ("flame0 640x480 q10", resize(f0, 640, 480, scale=240, quality=10, temporal_samples=1))] + \
                 [(n, resize(x, 256, 256, quality=50, temporal_samples=1)) for n, x in cases[1:6]]:
    gr, gg = ref.parse(xml), gpu.parse(xml)
    a = ref.render(gr, isaac_seed="a", nthreads=8)[0]; b = ref.render(gr, isaac_seed="b", nthreads=8)[0]
    res = []
    for wu in ("0", "1"):
        os.environ["F3B_WU"] = wu
        ps = []
        for seed in ("s1", "s2", "s3"):
            im = gpu.render(gg, isaac_seed=seed)[0]
            ps.append(min(psnr(im, a), psnr(im, b)))
        res.append(np.mean(ps))
    print("%-20s ref-vs-ref %.2f dB   gpu per-lane %.2f dB   gpu warp-uniform %.2f dB" % (name, psnr(a, b), res[0], res[1]), flush=True)
