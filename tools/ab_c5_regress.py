"""Which round-2 change costs the C5 genomes their speed: sweepers, aggregation (now with
integer atomics), or the 72-register budget? (run under gpurun)"""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
names = sys.argv[1:] or ["var96", "var24", "var64", "var97", "var47", "var58", "var76", "mix105", "var00", "var30"]
combos = [("default", {}), ("no sweep", {"F3B_SWEEP": "0"}), ("AGG=0", {"F3B_AGG": "0"}), ("AGG=1", {"F3B_AGG": "1"}),
          ("no sweep, AGG=0", {"F3B_SWEEP": "0", "F3B_AGG": "0"})]
for n in names:
    p = sorted(glob.glob(os.path.join(ROOT, "tests/golden/c5", "*%s.flam3" % n)))[0]
    g = gpu.parse(open(p).read())
    row = []
    for tag, env in combos:
        for k in ("F3B_SWEEP", "F3B_AGG"):
            os.environ.pop(k, None)
        os.environ.update(env)
        gpu.render(g, isaac_seed="perf"); gpu.render(g, isaac_seed="perf"); ctx.stats_reset()
        for _ in range(2):
            img, st, _ = gpu.render(g, isaac_seed="perf")
        it, nl = ctx.stage_stats(0)
        row.append("%s %.3e (%d launches)" % (tag, st["num_iters"] / (it / 2 / 1e3), nl // 2))
    print("%-8s " % n + " | ".join(row), flush=True)
