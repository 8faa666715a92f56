"""C3-shaped frames on one GPU: binned scatter against the direct scatter into a
DRAM-resident histogram (run under gpurun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
def run(name, xml, env, reps=2, **kw):
    for k in ("F3B_BINNED", "F3B_BIN_REGION_MB", "F3B_BIN_MEM_MB"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = gpu.parse(xml)
    gpu.render(g, **kw)
    ctx.stats_reset()
    t = time.perf_counter()
    for _ in range(reps):
        img, st, _ = gpu.render(g, **kw)
    wall = (time.perf_counter() - t) / reps
    it, n1 = ctx.stage_stats(0); de, _ = ctx.stage_stats(1); fi, _ = ctx.stage_stats(2)
    print("%-44s wall %8.2f ms | iterate %8.3f (%d launches) density %7.3f final %7.3f | %.3e it/s (stage) lost %s mean %.1f"
          % (name, wall * 1e3, it / reps, n1 / reps, de / reps, fi / reps, st["num_iters"] / (it / reps / 1e3), ctx.exactness(), img.mean()), flush=True)
def c3(q, ts):
    return "<flames>" + resize(f0, 3840, 2160, scale=1080, quality=q, oversample=2, temporal_samples=ts) + \
      resize(f1, 3840, 2160, scale=1080, quality=q, oversample=2, temporal_samples=ts).replace('time="0"', 'time="100"', 1) + "</flames>"
kw = dict(ngenomes=2, time=50.0)
run("C3 q500 ts=1000 direct (DRAM scatter)", c3(500, 1000), {"F3B_BINNED": "0"}, **kw)
run("C3 q500 ts=1000 binned 48 MB regions", c3(500, 1000), {}, **kw)
run("C3 q500 ts=1000 binned 32 MB regions", c3(500, 1000), {"F3B_BIN_REGION_MB": "32"}, **kw)
run("C3 q500 ts=1000 binned 64 MB regions", c3(500, 1000), {"F3B_BIN_REGION_MB": "64"}, **kw)
run("C3 q500 ts=1000 binned 48 MB, 4 GB pool", c3(500, 1000), {"F3B_BIN_MEM_MB": "4096"}, **kw)
run("C3 q2000 ts=1000 binned 48 MB", c3(2000, 1000), {}, reps=1, **kw)
run("C3 q2000 ts=1000 direct", c3(2000, 1000), {"F3B_BINNED": "0"}, reps=1, **kw)
still = resize(f0, 3840, 2160, scale=1080, quality=500, oversample=2, temporal_samples=1)
run("4K os2 still q500 direct", still, {"F3B_BINNED": "0"})
run("4K os2 still q500 binned", still, {})
