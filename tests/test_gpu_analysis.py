"""the analysis helpers either side of the render, on the GPU.
flam3_estimate_bounding_box (flam3.c:3574-3639) and flam3_colorhist (flam3.c:351-394) are
flam3_iterate plus host arithmetic; the drop-in runs the iterations as many independent
chains on the GPU (f3b_orbits). Their results are statistics of the attractor, so they must
agree with the unmodified reference within the reference's own seed-to-seed spread."""
import glob
import os

import numpy as np
import pytest

from parity_util import libs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _genomes():
    out = [("flame0", open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()),
           ("flame1", open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read())]
    for name in ("mix101", "mix107", "var13", "var33"):
        p = sorted(glob.glob(os.path.join(ROOT, "tests/golden/c5", "*%s.flam3" % name)))
        if p:
            out.append((name, open(p[0]).read()))
    return out


@pytest.mark.parametrize("name,xml", _genomes())
def test_bounding_box_matches_the_reference(name, xml):
    ref, gpu, _ = libs()
    gr, gg = ref.parse(xml), gpu.parse(xml)
    boxes = [np.array(ref.bounding_box(gr, eps=0.01, nsamples=100000, seed=s)[1]) for s in ("a", "b", "c")]
    rv, g = gpu.bounding_box(gg, eps=0.01, nsamples=100000, seed="a")
    assert rv >= 0
    g = np.array(g)
    size = max(boxes[0][2] - boxes[0][0], boxes[0][3] - boxes[0][1])
    spread = max(np.abs(boxes[0] - boxes[1]).max(), np.abs(boxes[0] - boxes[2]).max())
    err = min(np.abs(g - b).max() for b in boxes)
    # as close to a reference box as the reference's boxes are to each other, plus 2 % of the size
    assert err <= 2.0 * spread + 0.02 * size, (name, g, boxes, spread)
    assert g[0] < g[2] and g[1] < g[3]


@pytest.mark.parametrize("name,xml", _genomes()[:4])
def test_colorhist_matches_the_reference(name, xml):
    ref, gpu, _ = libs()
    gr, gg = ref.parse(xml), gpu.parse(xml)
    ra = ref.colorhist(gr, num_batches=20, seed="a")[1]
    rb = ref.colorhist(gr, num_batches=20, seed="b")[1]
    rv, g = gpu.colorhist(gg, num_batches=20, seed="a")
    assert rv == 0
    assert abs(g.sum() - 1.0) < 1e-6 and abs(ra.sum() - 1.0) < 1e-6
    noise = np.abs(ra - rb).sum()
    assert min(np.abs(g - ra).sum(), np.abs(g - rb).sum()) <= 1.5 * noise + 0.02, (name, noise)
