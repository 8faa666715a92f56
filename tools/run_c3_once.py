"""One C3-shaped render (for ncu launch lists): env F3B_* select the path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu = fb.load_dropin()
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
q = float(sys.argv[1]) if len(sys.argv) > 1 else 500.0
ts = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
xml = "<flames>" + resize(f0, 3840, 2160, scale=1080, quality=q, oversample=2, temporal_samples=ts) + \
      resize(f1, 3840, 2160, scale=1080, quality=q, oversample=2, temporal_samples=ts).replace('time="0"', 'time="100"', 1) + "</flames>"
g = gpu.parse(xml)
for _ in range(reps):
    img, st, _ = gpu.render(g, ngenomes=2, time=50.0)
print(st, img.mean())
