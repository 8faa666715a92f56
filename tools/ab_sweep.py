"""A/B on one box: cost of the sweeper warps (exact integer buckets) on the C2 chaos game,
and what they buy: buckets whose sums pass 2^24 (f3b_iterate.cu)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
g = gpu.parse(resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1))
geo = gpu.buffer_geometry(g)
ctx.width, ctx.height = geo["width"], geo["height"]
def t(env, reps=5):
    for k in ("F3B_SWEEP", "F3B_SWEEP_GAP_US", "F3B_SWEEP_MODE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    gpu.render(g, isaac_seed="ab"); ctx.stats_reset()
    for _ in range(reps): gpu.render(g, isaac_seed="ab")
    ms = ctx.stage_stats(0)[0] / reps
    import ctypes as C
    out = (C.c_uint64 * 16)()
    core._dll.f3b_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    core._dll.f3b_debug_counters(ctx.ptr, out)
    return ms, ctx.exactness() + ("lost seen by cold/hot/merge", out[10], out[11], out[12], "hot entries", out[8] >> 32)
hist = {}
for rnd in range(2):
    for name, env in (("off", {"F3B_SWEEP": "0"}), ("auto", {}), ("norep", {"F3B_SWEEP_MODE": "3"}), ("flat", {"F3B_SWEEP_GAP_US": "1"}),
                      ("4ms", {"F3B_SWEEP_GAP_US": "4000"}), ("idle", {"F3B_SWEEP_MODE": "0"}),
                      ("cold", {"F3B_SWEEP_MODE": "1"}), ("hot", {"F3B_SWEEP_MODE": "2"})):
        ms, ex = t(env)
        print("round %d: sweep %-5s chaos game %.3f ms  lost/saturated %s" % (rnd, name, ms, ex), flush=True)
        if rnd == 0 and name in ("off", "flat"):
            hist[name] = ctx.hist_download_u32().astype(np.int64)
a, b = hist["off"], hist["flat"]
d = (a != b).any(axis=-1)
print("fullest bucket count %d (2^24 = %d); buckets differing between staged-only and swept: %d; "
      "of those below 2^24 in every channel: %d" % (b[..., 4].max(), 2 ** 24, int(d.sum()),
      int((d & (b.max(axis=-1) < 2 ** 24)).sum())))
hot = np.argsort(b[..., 4].ravel())[-5:]
for i in hot:
    print("  bucket %d: swept %s  staged-only %s" % (i, b.reshape(-1, 5)[i], a.reshape(-1, 5)[i]))
