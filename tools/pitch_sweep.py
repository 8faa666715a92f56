"""Chaos-game time of config C2 as a function of the histogram row pitch (F3B_HIST_PITCH):
identical work and image, only the address -> L2 slice mapping changes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
H = W * 9 // 16
g = gpu.parse(resize(f0, W, H, scale=W * 540.0 / 1920, quality=2000 * (1920.0 / W) ** 2, oversample=1, temporal_samples=1))
width = W + 18
pitches = sorted(set(list(range(width, width + 40, 2)) + list(range(width + 40, width + 200, 8)) +
                     [1 << k for k in range(8, 14) if (1 << k) >= width and (1 << k) < 3 * width] +
                     [((width + 127) // 128) * 128, ((width + 255) // 256) * 256, ((width + 511) // 512) * 512,
                      ((width + 63) // 64) * 64, ((width + 127) // 128) * 128 + 8, ((width + 127) // 128) * 128 + 32]))
ref_img = None
for p in pitches:
    os.environ["F3B_HIST_PITCH"] = str(p)
    img, st, _ = gpu.render(g, isaac_seed="pitch"); ctx.stats_reset()
    if ref_img is None: ref_img = img
    same = bool((img == ref_img).all())
    for _ in range(3): gpu.render(g, isaac_seed="pitch")
    ms = ctx.stage_stats(0)[0] / 3
    print("pitch %5d buckets (%6d B, %5.1f x128B): iterate %7.3f ms  %.3e it/s  image_same=%s" % (
        p, p * 16, p * 16 / 128.0, ms, st["num_iters"] / ms * 1e3, same), flush=True)
