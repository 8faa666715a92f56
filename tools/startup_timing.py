"""Where does the first frame's second go? dlopen, context creation, first and second render."""
import os, sys, time
t0 = time.perf_counter()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import set_flame_attrs
t1 = time.perf_counter()
core = fb.load_core(); gpu = fb.load_dropin()
t2 = time.perf_counter()
ctx = gpu.render_context(core)          # f3b_ctx_create: cudaSetDevice, stream, occupancy queries
t3 = time.perf_counter()
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
g = gpu.parse(set_flame_attrs(f0, temporal_samples=1))
t4 = time.perf_counter()
gpu.render(g)
t5 = time.perf_counter()
gpu.render(g)
t6 = time.perf_counter()
print("python imports %.0f ms | dlopen libs %.0f ms | context %.0f ms | parse %.1f ms | first render %.0f ms | second render %.1f ms"
      % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t6 - t5)))
