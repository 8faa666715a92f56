"""Render one genome file a few times (for ncu captures): run_genome.py <file.flam3> [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb
gpu, core = fb.load_dropin(), fb.load_core()
g = gpu.parse(open(sys.argv[1]).read())
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    img, st, _ = gpu.render(g, isaac_seed="perf")
print(st, img.mean())
