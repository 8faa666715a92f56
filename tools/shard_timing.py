import os, sys, time
sys.path.insert(0, "/root/repo")
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu, core = fb.load_dropin(), fb.load_core()
ctx = gpu.render_context(core)
f0 = open("/root/repo/tests/golden/test_flame0.flam3").read()
def run(name, q, shard=(0,1), reps=5, **kw):
    g = gpu.parse(resize(f0, 1920, 1080, scale=540, quality=q, oversample=1, temporal_samples=1))
    gpu.set_shard(*shard)
    gpu.render(g, **kw); gpu.render(g, **kw)
    ctx.stats_reset()
    for _ in range(reps):
        img, st, _ = gpu.render(g, **kw)
    it, n = ctx.stage_stats(0)
    print("%-30s iterate %.3f ms" % (name, it / reps), flush=True)
    gpu.set_shard(0, 1)
run("q2000 full", 2000)
run("q2000 shard 0/2", 2000, (0, 2))
run("q2000 shard 1/2", 2000, (1, 2))
run("q1000 full", 1000)
run("q500 full", 500)
run("q4000 full", 4000)
run("q2000 shard 0/4", 2000, (0, 4))
run("q2000 full again", 2000)
run("q2000 full sub_batch 5000", 2000, sub_batch_size=5000)
run("q2000 full sub_batch 20000", 2000, sub_batch_size=20000)
run("q2000 shard 0/8", 2000, (0, 8))
run("q2000 shard 3/8", 2000, (3, 8))
