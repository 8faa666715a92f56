"""The named configurations of BASELINE.json at (or close to) their real sizes, both parity
gates (north_star): C2's 1938x1098 bucket grid, C3's 4K oversample-2 motion-blurred frame
(7718x4358 buckets), C5's random genomes at 1024x1024, and the fullest buckets of the headline
frame, whose sums are 100x beyond what a float holds exactly.

The reference side is the UNMODIFIED reference (oracle/_ref); gate 1 runs it single-threaded
(its DE scatter races across threads) at a low quality so that it finishes in
seconds to a minute: the post-pass work does not depend on quality, only the buckets' values do.
"""
import glob
import os

import numpy as np
import pytest

from flam3_b200.genome_xml import resize
from parity_util import gpu_render_from_histogram, libs, psnr, ref_render_capture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(rimg, gimg, lsb=1, flipped=0.02):
    diff = np.abs(gimg.astype(np.int64) - rimg.astype(np.int64))
    assert diff.max() <= lsb, "max diff %d" % diff.max()
    assert (diff > 0).mean() <= flipped, "%.4f of bytes differ" % (diff > 0).mean()
    assert rimg.mean() > 1.0


def _strips_tile_whole(gpu, core, gg, buckets, geo, whole, nranks, **kw):
    from flam3_b200.multigpu import strip_rows

    H = whole.shape[0]
    tiled = np.zeros_like(whole)
    try:
        gpu.set_exchange(lambda ctx: None)
        gpu.set_strip_gather(lambda ctx, image, row_bytes, rows_per: None)
        for r in range(nranks):
            gpu.set_shard(r, nranks)
            img, _ = gpu_render_from_histogram(gpu, core, gg, buckets, geo, **kw)
            r0, r1, _ = strip_rows(H, r, nranks)
            tiled[r0:r1] = img[r0:r1]
    finally:
        gpu.set_shard(0, 1)
        gpu.set_exchange(None)
        gpu.set_strip_gather(None)
    assert np.array_equal(tiled, whole), "%d ranks: %d bytes differ" % (nranks, int((tiled != whole).sum()))


def test_c2_grid_gate1_and_strips(flame0):
    """Gate 1 on C2's own bucket grid (1938x1098: tile, halo and row-range arithmetic at the
    size the headline number is quoted on), whole frame and as 3 and 8 row strips."""
    ref, gpu, core = libs()
    xml = resize(flame0, 1920, 1080, scale=540, quality=20, oversample=1, estimator_radius=9,
                 temporal_samples=1)
    gr, gg = ref.parse(xml), gpu.parse(xml)
    rimg, _, buckets, _, geo, _ = ref_render_capture(ref, gr)
    assert (geo["width"], geo["height"]) == (1938, 1098)
    gimg, _ = gpu_render_from_histogram(gpu, core, gg, buckets, geo)
    _check(rimg, gimg)
    for nranks in (3, 8):
        _strips_tile_whole(gpu, core, gg, buckets, geo, gimg, nranks)


def _c3_xml(flame0, flame1, w, h, scale, quality, ts):
    a = resize(flame0, w, h, scale=scale, quality=quality, oversample=2, estimator_radius=9, temporal_samples=ts)
    b = resize(flame1, w, h, scale=scale, quality=quality, oversample=2, estimator_radius=9, temporal_samples=ts)
    return "<flames>" + a + b.replace('time="0"', 'time="100"', 1) + "</flames>"


def test_c3_full_size_gate1(flame0, flame1):
    """C3's frame - 3840x2160, oversample 2, two key frames rendered at time 50 with temporal
    samples (motion blur) - at its real size: 7718x4358 buckets, filter width 8, DE kernels up
    to radius 18. The reference's own buckets (672 MB) through both post-passes, <= 1 LSB; the
    8-rank strips tile the frame."""
    ref, gpu, core = libs()
    xml = _c3_xml(flame0, flame1, 3840, 2160, 1080, 0.5, 100)
    kw = dict(ngenomes=2, time=50.0)
    gr, gg = ref.parse(xml), gpu.parse(xml)
    rimg, rst, buckets, _, geo, _ = ref_render_capture(ref, gr, **kw)
    assert (geo["width"], geo["height"], geo["filter_width"]) == (7718, 4358, 8)
    gimg, _ = gpu_render_from_histogram(gpu, core, gg, buckets, geo, **kw)
    _check(rimg, gimg)
    _strips_tile_whole(gpu, core, gg, buckets, geo, gimg, 8, **kw)


def test_c3_shape_gate2(flame0, flame1):
    """C3's structure at a quarter of the linear size, full pipeline on both sides: oversample
    2, two key frames at time 50, 100 temporal samples."""
    ref, gpu, _ = libs()
    xml = _c3_xml(flame0, flame1, 960, 540, 270, 60, 100)
    kw = dict(ngenomes=2, time=50.0)
    gr, gg = ref.parse(xml), gpu.parse(xml)
    a, sa, _ = ref.render(gr, isaac_seed="c3-a", nthreads=8, **kw)
    b, _, _ = ref.render(gr, isaac_seed="c3-b", nthreads=8, **kw)
    g, sg, _ = gpu.render(gg, isaac_seed="c3-a", nthreads=8, **kw)
    assert sa["rv"] == 0 and sg["rv"] == 0 and sg["num_iters"] == sa["num_iters"]
    assert min(psnr(g, a), psnr(g, b)) >= psnr(a, b) - 1.5, (psnr(g, a), psnr(g, b), psnr(a, b))
    assert abs(g.mean() - a.mean()) <= max(0.5, 0.02 * a.mean())


C5_SUBSET = ["mix107", "mix116", "mix123", "mix101", "var76", "var58", "var47", "var33", "var13", "var50",
             "var98", "var00"]


@pytest.mark.parametrize("name", C5_SUBSET)
def test_c5_at_1024(name):
    """C5 at the configuration's own frame size (1024x1024, 1042x1042 buckets), quality 50 on
    both sides: the heaviest mixed genomes, the ones whose attractors collapse (var76/58/47) and
    a spread of single-variation genomes."""
    ref, gpu, _ = libs()
    paths = glob.glob(os.path.join(ROOT, "tests", "golden", "c5", "*%s.flam3" % name)) + \
        glob.glob(os.path.join(ROOT, "tests", "golden", "c5", "*%s_*.flam3" % name))
    assert paths, name
    xml = resize(open(sorted(paths)[0]).read(), 1024, 1024, quality=50, temporal_samples=1)
    gr, gg = ref.parse(xml), gpu.parse(xml)
    a, sa, _ = ref.render(gr, isaac_seed="c5-a", nthreads=8)
    b, _, _ = ref.render(gr, isaac_seed="c5-b", nthreads=8)
    g, sg, _ = gpu.render(gg, isaac_seed="c5-a", nthreads=8)
    assert sa["rv"] == 0 and sg["rv"] == 0 and sg["num_iters"] == sa["num_iters"]
    # a collapsed attractor gives (nearly) the same image whatever the seed: beyond 60 dB the
    # images are equal for every purpose (psnr() itself reports identical images as 99)
    noise, got = min(psnr(a, b), 60.0), min(psnr(g, a), psnr(g, b))
    assert got >= noise - 1.5, "%s: gpu-vs-ref %.2f dB, ref-vs-ref %.2f dB" % (name, got, noise)


def test_hot_buckets_of_the_headline_frame(flame0):
    """The 1000 fullest buckets of C2 (quality 2000 on the GPU): their colour-to-count ratios
    r/a, g/a, b/a are the mean palette colour of the hits that land there - a property of the
    attractor, not of the quality - so they must agree with the reference's buckets (quality
    200) within the reference's sampling noise. Float buckets drifted by ~1 % here (their sums
    pass 2^24 a hundredfold); the integer buckets must not."""
    ref, gpu, core = libs()
    gg = gpu.parse(resize(flame0, 1920, 1080, scale=540, quality=2000, oversample=1, estimator_radius=9,
                          temporal_samples=1))
    gr = ref.parse(resize(flame0, 1920, 1080, scale=540, quality=200, oversample=1, estimator_radius=9,
                          temporal_samples=1))
    _, _, rb, _, geo, _ = ref_render_capture(ref, gr, nthreads=8)
    _, st, _ = gpu.render(gg, isaac_seed="hot")
    ctx = gpu.render_context(core)
    ctx.width, ctx.height = geo["width"], geo["height"]
    gb = ctx.hist_download_u32().astype(np.float64).reshape(-1, 5)
    lost, _ = ctx.exactness()
    assert lost == 0
    rb = rb.astype(np.float64).reshape(-1, 5)
    top = np.argsort(gb[:, 4])[-1000:]
    assert gb[top, 4].min() > 2 ** 24            # all of them beyond float exactness
    n_ref = rb[top, 4] / 255.0                   # hits behind the reference's ratios
    assert n_ref.min() > 1000
    worst = 0.0
    for ch in range(3):
        rg, rr = gb[top, ch] / gb[top, 3], rb[top, ch] / rb[top, 3]
        # a palette channel is in [0, 1]: its standard deviation is at most 0.5
        z = np.abs(rg - rr) / (0.5 / np.sqrt(n_ref))
        worst = max(worst, float(z.max()))
        assert z.max() < 6.0, "channel %d: %.1f sigma" % (ch, z.max())
        # and no common bias: the mean deviation over 1000 buckets is far below one sigma
        assert abs(np.mean((rg - rr) / (0.5 / np.sqrt(n_ref)))) < 0.5
    # hit shares of those buckets agree too (same attractor density)
    share_g = gb[top, 4].sum() / gb[:, 4].sum()
    share_r = rb[top, 4].sum() / rb[:, 4].sum()
    assert abs(share_g - share_r) < 0.02 * share_r, (share_g, share_r)
