"""Regression tests for defects found in review."""
import os

import numpy as np
import pytest

from flam3_b200.genome_xml import resize, set_flame_attrs
from parity_util import libs, postpass_parity, psnr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("estimator,oversample", [(9, 4), (10, 4), (20, 2), (9, 5)])
def test_wide_density_estimation_kernels(flame0, estimator, oversample):
    """Kernel radius R = estimator*oversample: the staged gather holds (32+2R)^2 buckets in
    shared memory, which stops fitting at R = 38; wider kernels go through the unstaged gather
    (f3b_post.cu). (9,4) is the widest staged case, the others are beyond it. Gate 1: the
    reference's own histogram through both post-passes, <= 1 LSB."""
    xml = resize(flame0, 96, 72, scale=36, quality=5, oversample=oversample,
                 estimator_radius=estimator, temporal_samples=1)
    maxdiff, frac = postpass_parity(xml)
    assert maxdiff <= 1 and frac < 0.02, (maxdiff, frac)


TRAILING_ZERO = """
<flame time="0" size="128 128" center="0 0" scale="32" quality="60" temporal_samples="1"
       estimator_radius="0" brightness="4" gamma="4" vibrancy="1" supersample="1" filter="0.5">
   <xform weight="0.5" color="0.1" color_speed="1" linear="1" coefs="0.5 0 0 0.5 -0.5 -0.5"/>
   <xform weight="0.5" color="0.5" color_speed="1" linear="1" coefs="0.5 0 0 0.5 0.5 0.5"/>
   <xform weight="0" color="0.9" color_speed="1" linear="1" coefs="0.5 0 0 0.5 0 0.9"/>
   %s
</flame>
"""


def test_zero_weight_trailing_xform_is_never_picked():
    """The packed pick table of the SIMPLE kernel sent k = 16383 to the LAST xform even when
    its weight is 0 (the reference's table holds an earlier xform there, flam3.c:196-216):
    one iteration in 16384 stamped the third xform's colour. Pure colours per xform make any
    such hit visible in the histogram totals; SIMPLE and generic kernels must agree."""
    _, gpu, core = libs()
    pal = []
    for i in range(256):
        c = (255, 0, 0) if i < 85 else ((0, 255, 0) if i < 170 else (0, 0, 255))
        pal.append('<color index="%d" rgb="%d %d %d"/>' % ((i,) + c))
    gg = gpu.parse(TRAILING_ZERO % "\n".join(pal))
    geo = gpu.buffer_geometry(gg)
    ctx = gpu.render_context(core)
    ctx.width, ctx.height = geo["width"], geo["height"]
    hists = {}
    try:
        for mode in ("simple", "generic"):
            if mode == "generic":
                os.environ["F3B_NO_SIMPLE"] = "1"
            _, st, _ = gpu.render(gg, isaac_seed="zw")
            assert st["rv"] == 0
            hists[mode] = ctx.hist_download_u32().astype(np.int64)
    finally:
        os.environ.pop("F3B_NO_SIMPLE", None)
    for mode, h in hists.items():
        tot = h[..., :3].reshape(-1, 3).sum(axis=0)
        assert tot[0] > 0 and tot[1] > 0
        # color_speed = 1: a hit carries exactly its xform's colour; a 1/16384 leak would be
        # ~60 hits of pure blue here
        assert tot[2] == 0, (mode, tot)
    assert np.array_equal(hists["simple"], hists["generic"])
