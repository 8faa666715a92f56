"""Chaos-game throughput on the random genomes of config C5 at 1024x1024 quality 1000 (run
under gpurun; writes gpurun_out/c5_perf.txt)."""
import glob, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
rows = []
for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "c5", "c5_*.flam3"))):
    xml = open(p).read()
    g = gpu.parse(xml)
    gpu.render(g, isaac_seed="perf")
    ctx.stats_reset()
    reps = 2
    for _ in range(reps):
        img, st, _ = gpu.render(g, isaac_seed="perf")
    it, _ = ctx.stage_stats(0)
    de, _ = ctx.stage_stats(1)
    nx = len(re.findall(r"<xform", xml))
    rows.append((os.path.basename(p)[3:-6], nx, 1 if "<finalxform" in xml else 0,
                 st["num_iters"] / (it / reps / 1e3), it / reps, de / reps, st["badvals"], img.mean()))
rows.sort(key=lambda r: r[3])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "c5_perf.txt"), "w") as f:
    f.write("# genome xforms final iters_per_s iterate_ms density_ms badvals image_mean   (1024x1024 q1000, 1.05e9 iterations)\n")
    for r in rows:
        line = "%-8s %2d %d %.3e %8.3f %7.3f %10d %6.1f" % r
        f.write(line + "\n")
        print(line)
    import statistics
    v = [r[3] for r in rows]
    f.write("# median %.3e  min %.3e  max %.3e iters/s over %d genomes\n" % (statistics.median(v), min(v), max(v), len(v)))
    print("median %.3e min %.3e max %.3e" % (statistics.median(v), min(v), max(v)))
