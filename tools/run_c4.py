"""Render config C4's frame (test.flam3's two flames at time 50, 1000 temporal samples)
a few times with F3B_TIMING host-phase output."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["F3B_TIMING"] = "1"
import flam3_b200 as fb
gpu, core = fb.load_dropin(), fb.load_core()
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
g = gpu.parse("<flames>" + f0 + f1 + "</flames>")
for _ in range(4):
    gpu.render(g, ngenomes=2, time=50.0, nthreads=16)  # flam3-animate passes the core count
