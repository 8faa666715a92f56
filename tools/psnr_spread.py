"""How much does the small-image PSNR statistic of tests/test_gpu_variations.py scatter with
the seed? Per variation: 6 reference renders (15 pairs), 4 GPU renders (24 GPU-reference pairs):
range, mean and median of both sets (run under gpurun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import libs, psnr
from variation_cases import VAR_NAMES, genome_xml
ref, gpu, core = libs()
worst = 0.0
for vid in [int(a) for a in sys.argv[1:]] or range(99):
    xml = genome_xml(vid, size=96, quality=40)
    gr, gg = ref.parse(xml), gpu.parse(xml)
    refs = [ref.render(gr, isaac_seed="v%s" % "abcdef"[i], nthreads=1)[0] for i in range(6)]
    rr = [min(psnr(refs[i], refs[j]), 60.0) for i in range(6) for j in range(i + 1, 6)]
    gs = [gpu.render(gg, isaac_seed="v%s" % "abcd"[i])[0] for i in range(4)]
    gv = [min(psnr(g, r), 60.0) for g in gs for r in refs]
    d = np.median(gv) - np.median(rr)
    worst = min(worst, d)
    print("%2d %-16s ref-vs-ref %.2f..%.2f mean %.2f median %.2f   gpu-vs-ref %.2f..%.2f mean %.2f median %.2f   median diff %+.2f  mean diff %+.2f"
          % (vid, VAR_NAMES[vid], min(rr), max(rr), np.mean(rr), np.median(rr), min(gv), max(gv), np.mean(gv), np.median(gv), d,
             np.mean(gv) - np.mean(rr)), flush=True)
print("worst median diff %.2f" % worst)
