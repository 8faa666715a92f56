"""First contact with the GPU: render fixtures with the drop-in and the reference."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import flam3_b200 as fb
from flam3_b200.genome_xml import set_flame_attrs, resize
from imgio import write_png

def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)

out = os.path.join(ROOT, "gpurun_out"); os.makedirs(out, exist_ok=True)
gpu = fb.load_dropin()
ref = fb.Flam3Lib(os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so"))
xml = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
xml1 = set_flame_attrs(xml, temporal_samples=1)
for name, x, kw in [("q10", xml1, {}), ("q100", set_flame_attrs(xml1, quality=100), {}),
                    ("q10_ts1000", xml, {})]:
    gg, gr = gpu.parse(x), ref.parse(x)
    t = time.time(); ig, sg, _ = gpu.render(gg); tg = time.time() - t
    t = time.time(); ir, sr, _ = ref.render(gr, nthreads=os.cpu_count()); tr = time.time() - t
    print(name, "gpu", sg, "%.3fs" % tg, "ref", sr, "%.3fs" % tr, "psnr %.2f" % psnr(ig, ir),
          "means", ig.mean(), ir.mean(), flush=True)
    write_png(os.path.join(out, "fl_%s_gpu.png" % name), ig)
    write_png(os.path.join(out, "fl_%s_ref.png" % name), ir)
# C2: 1080p q2000
x = resize(xml1, 1920, 1080, scale=540, quality=2000, oversample=1, estimator_radius=9)
gg = gpu.parse(x)
for i in range(3):
    t = time.time(); ig, sg, _ = gpu.render(gg); tg = time.time() - t
    print("C2 gpu", sg, "%.3fs" % tg, "iters/s %.3e" % (sg["num_iters"] / tg), flush=True)
write_png(os.path.join(out, "fl_c2_gpu.png"), ig)
x = resize(xml1, 1920, 1080, scale=540, quality=50, oversample=1, estimator_radius=9)
gr = ref.parse(x)
t = time.time(); ir, sr, _ = ref.render(gr, nthreads=os.cpu_count()); tr = time.time() - t
print("C2@q50 ref", sr, "%.3fs" % tr, "iters/s %.3e" % (sr["num_iters"] / tr), "cores", os.cpu_count())
gg = gpu.parse(x); ig, sg, _ = gpu.render(gg)
print("C2@q50 psnr %.2f" % psnr(ig, ir))
write_png(os.path.join(out, "fl_c2q50_ref.png"), ir); write_png(os.path.join(out, "fl_c2q50_gpu.png"), ig)
