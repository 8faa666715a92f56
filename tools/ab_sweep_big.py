"""Sweepers on histograms that live in DRAM (4K oversample-2 still, 7718x4358 buckets, 538 MB of
staging): how many cold passes a frame gets, what the hot list holds, and whether any staged
value was seen at 2^24 (run under gpurun)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
core._dll.f3b_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
def run(name, xml, env, reps=2):
    for k in ("F3B_SWEEP", "F3B_SWEEP_GAP_US", "F3B_SWEEP_MODE", "F3B_BINNED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = gpu.parse(xml)
    geo = gpu.buffer_geometry(g)
    gpu.render(g, isaac_seed="big"); ctx.stats_reset()
    for _ in range(reps): _, st, _ = gpu.render(g, isaac_seed="big")
    ms, nl = ctx.stage_stats(0)
    out = (C.c_uint64 * 16)()
    core._dll.f3b_debug_counters(ctx.ptr, out)
    nseg = (geo["width"] * geo["height"] + 255) // 256
    print("%-40s iterate %8.3f ms (%d launches)  lost/sat %s  by cold/hot/merge %d %d %d  hot entries %d  "
          "cold tickets of the last launch %d = %.1f passes" % (name, ms / reps, nl / reps, ctx.exactness(), out[10], out[11], out[12],
          out[8] >> 32, out[8] & 0xffffffff, (out[8] & 0xffffffff) / nseg), flush=True)
for q in (500, 2000):
    still = resize(f0, 3840, 2160, scale=1080, quality=q, oversample=2, temporal_samples=1)
    for name, env in (("cold passes + hot list only", {"F3B_SWEEP_MODE": "3"}), ("+ walkers' reports (default)", {}),
                      ("no sweepers", {"F3B_SWEEP": "0"})):
        run("4K os2 still q%d %s" % (q, name), still, dict(env, F3B_BINNED="0"), reps=2 if q == 500 else 1)
for w, h, os_ in ((3840, 2160, 1), (7680, 4320, 1), (7680, 4320, 2)):
    still = resize(f0, w, h, scale=h / 2, quality=500 if os_ == 1 else 125, oversample=os_, temporal_samples=1)
    for name, env in (("cold + hot only", {"F3B_SWEEP_MODE": "3"}), ("default", {})):
        run("%dx%d os%d still %s" % (w, h, os_, name), still, dict(env, F3B_BINNED="0"))
