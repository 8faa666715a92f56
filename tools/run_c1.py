"""Render config C1 (test.flam3 flame 0 as shipped, one temporal sample) a few times — the
latency-bound regime: 308 chains of 10 000 sequential iterations. For ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import set_flame_attrs
gpu, core = fb.load_dropin(), fb.load_core()
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
g = gpu.parse(set_flame_attrs(f0, temporal_samples=1))
for _ in range(4):
    gpu.render(g)
