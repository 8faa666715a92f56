"""How much does the small-image PSNR statistic of tests/test_gpu_variations.py scatter with
the seed? For one variation: reference-vs-reference over several seed pairs, GPU-vs-reference
over several GPU seeds (run under gpurun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import libs, psnr
from variation_cases import VAR_NAMES, genome_xml
ref, gpu, core = libs()
for vid in [int(a) for a in sys.argv[1:]] or [27]:
    xml = genome_xml(vid, size=96, quality=40)
    gr, gg = ref.parse(xml), gpu.parse(xml)
    refs = [ref.render(gr, isaac_seed="r%d" % i, nthreads=1)[0] for i in range(6)]
    rr = [psnr(refs[i], refs[j]) for i in range(6) for j in range(i + 1, 6)]
    for wu in ("0", "default"):
        if wu == "0": os.environ["F3B_WU"] = "0"
        else: os.environ.pop("F3B_WU", None)
        gs = [gpu.render(gg, isaac_seed="g%d" % i)[0] for i in range(8)]
        gv = [psnr(g, r) for g in gs for r in refs]
        gg_ = [psnr(gs[i], gs[j]) for i in range(8) for j in range(i + 1, 8)]
        print("%s WU=%s: ref-vs-ref %.2f..%.2f (mean %.2f)   gpu-vs-ref %.2f..%.2f (mean %.2f)   gpu-vs-gpu %.2f..%.2f (mean %.2f)"
              % (VAR_NAMES[vid], wu, min(rr), max(rr), np.mean(rr), min(gv), max(gv), np.mean(gv), min(gg_), max(gg_), np.mean(gg_)), flush=True)
