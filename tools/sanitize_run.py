"""Small renders that touch every kernel and mode (for compute-sanitizer):
SIMPLE / generic chaos game, count plane, xaos + final xform, DE with oversample, no DE,
earlyclip, 16-bit RGBA, 1000-sample animation frame (side streams), strips."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize, rescale, set_flame_attrs
gpu, core = fb.load_dropin(), fb.load_core()
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
def r(name, xml, **kw):
    g = gpu.parse(xml)
    img, st, _ = gpu.render(g, **kw)
    assert st["rv"] == 0, (name, st)
    print(name, img.shape, round(float(img.mean()), 2), flush=True)
r("simple", resize(f0, 160, 120, scale=60, quality=20, temporal_samples=1))
r("os2 earlyclip", resize(f0, 96, 72, scale=36, quality=20, oversample=2, temporal_samples=1), earlyclip=1)
r("os3 no DE", resize(f0, 64, 48, scale=24, quality=20, oversample=3, temporal_samples=1, estimator_radius=0))
r("bits32", resize(f0, 96, 72, scale=36, quality=20, temporal_samples=1), bits=32)
r("bits64", resize(f0, 96, 72, scale=36, quality=20, temporal_samples=1), bits=64)
r("rgba16", resize(f0, 96, 72, scale=36, quality=20, temporal_samples=1), nchan=4, transp=1, bytes_per_channel=2)
r("motion blur 1000 samples", "<flames>" + resize(f0, 96, 72, scale=36, quality=20) + resize(f1, 96, 72, scale=36, quality=20) + "</flames>",
  ngenomes=2, time=50.0, nthreads=4)
r("gaussian temporal filter (count plane)", "<flames>" + resize(f0, 96, 72, scale=36, quality=20, temporal_samples=40, temporal_filter_type="gaussian") +
  resize(f1, 96, 72, scale=36, quality=20, temporal_samples=40, temporal_filter_type="gaussian") + "</flames>", ngenomes=2, time=30.0)
for name in ("mix103", "mix105", "mix123", "var20", "var33", "var59"):
    r(name, rescale(open(os.path.join(ROOT, "tests/golden/c5/c5_%s.flam3" % name)).read(), 96, 96, quality=20))
# round 2 paths
def env(**kv):
    for k in ("F3B_CARRY", "F3B_SWEEP_MODE", "F3B_BINNED", "F3B_BIN_REGION_MB", "F3B_BIN_MEM_MB", "F3B_WU", "F3B_AGG"):
        os.environ.pop(k, None)
    os.environ.update(kv)
env(F3B_CARRY="600")
r("sweepers, carries at 600", resize(f0, 160, 120, scale=60, quality=100, temporal_samples=1))
env(F3B_CARRY="600", F3B_SWEEP_MODE="11")
r("sweepers + walkers' reports", resize(f0, 160, 120, scale=60, quality=100, temporal_samples=1))
r("... with a count plane", set_flame_attrs(resize(f0, 160, 120, scale=60, quality=100, temporal_samples=1), palette_mode="linear"))
env(F3B_BINNED="1", F3B_BIN_REGION_MB="0.25", F3B_BIN_MEM_MB="160")
r("binned scatter", resize(f0, 160, 120, scale=60, quality=40, temporal_samples=1), sub_batch_size=400)
env(F3B_WU="1")
r("warp-uniform step", rescale(open(os.path.join(ROOT, "tests/golden/c5/c5_mix107.flam3")).read(), 128, 128, quality=400), sub_batch_size=100)
env(F3B_AGG="1")
r("aggregated adds", resize(f0, 96, 72, scale=36, quality=100, temporal_samples=1))
env()
g = gpu.parse(resize(f0, 160, 120, scale=60, quality=20, temporal_samples=1))
print("bounding box (f3b_orbits)", gpu.bounding_box(g, 0, 0.01, 4000), flush=True)
print("colorhist (f3b_orbits)", gpu.colorhist(g, 0, 2)[0], flush=True)
gpu.set_shard(1, 3); gpu.set_exchange(lambda ctx: None); gpu.set_strip_gather(lambda c, i, rb, rp: None)
r("strip 1/3", resize(f0, 160, 120, scale=60, quality=20, temporal_samples=1))
print("done")
