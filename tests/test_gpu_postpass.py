"""Parity gate (1), BASELINE.json north_star: given an identical histogram dumped from the
reference, the GPU density estimation, spatial filter and tone map must reproduce the
reference's output image to <= 1 LSB (at 8 bpc; <= 1 LSB of the 8-bit scale, i.e. 256 steps,
at 16 bpc, where the reference itself quantises a*256).

The histogram and accumulator come from the UNMODIFIED reference (oracle/_ref, captured by
oracle/ref_wrap.c), run single-threaded so its DE scatter is deterministic.
Everything goes through the drop-in flam3_render, i.e. the real host path computes the
filters, gutters, k1/k2 and averaged tone-map parameters.
"""
import numpy as np
import pytest

from flam3_b200.genome_xml import resize, set_flame_attrs
from parity_util import gpu_render_from_histogram, libs, ref_render_capture

pytestmark = pytest.mark.gpu

# float accumulation order differs between the reference's scatter and the gather, so
# accumulators agree to float rounding, not bit for bit
ACC_RTOL = 2e-5
# fraction of output bytes allowed to sit on the other side of a truncation boundary
MAX_FLIPPED = 0.02


def run_case(xml, **kw):
    ref, gpu, core = libs()
    gr, gg = ref.parse(xml), gpu.parse(xml)
    rimg, rst, buckets, racc, geo, _ = ref_render_capture(ref, gr, **kw)
    gimg, gacc = gpu_render_from_histogram(gpu, core, gg, buckets, geo, **kw)
    return rimg, gimg, racc, gacc, geo, buckets


def check_images(rimg, gimg, lsb=1):
    diff = np.abs(gimg.astype(np.int64) - rimg.astype(np.int64))
    assert diff.max() <= lsb, "max diff %d" % diff.max()
    assert (diff > 0).mean() <= MAX_FLIPPED, "%.4f of bytes differ" % (diff > 0).mean()
    assert rimg.mean() > 1.0  # the case actually drew something


def small(xml, quality=20, **attrs):
    base = dict(temporal_samples=1)
    base.update(attrs)
    return resize(xml, 320, 240, scale=120, quality=quality, **base)


def test_default_de_os1(flame0):
    rimg, gimg, racc, gacc, geo, _ = run_case(small(flame0))
    check_images(rimg, gimg)
    # accumulators: compare where the reference wrote something meaningful
    scale = np.abs(racc).max()
    np.testing.assert_allclose(gacc, racc, rtol=ACC_RTOL, atol=ACC_RTOL * scale)


def test_native_test_flam3_both_flames(flame0, flame1):
    # config 1: test.flam3 at its native 640x480 q10
    for xml in (flame0, flame1):
        rimg, gimg, racc, gacc, geo, _ = run_case(set_flame_attrs(xml, temporal_samples=1))
        assert geo["width"] == 658 and geo["height"] == 498
        check_images(rimg, gimg)


def test_oversample2(flame0):
    rimg, gimg, *_ = run_case(small(flame0, supersample=2))
    check_images(rimg, gimg)


def test_oversample3_small_estimator(flame0):
    rimg, gimg, *_ = run_case(small(flame0, supersample=3, estimator_radius=3,
                                    estimator_minimum=1, estimator_curve=0.6))
    check_images(rimg, gimg)


def test_no_density_estimation(flame0):
    rimg, gimg, racc, gacc, *_ = run_case(small(flame0, estimator_radius=0))
    check_images(rimg, gimg)
    np.testing.assert_allclose(gacc, racc, rtol=ACC_RTOL, atol=1e-6)


def test_high_quality_small_kernels(flame0):
    rimg, gimg, *_ = run_case(small(flame0, quality=400))
    check_images(rimg, gimg)


def test_earlyclip(flame0):
    rimg, gimg, *_ = run_case(small(flame0), earlyclip=1)
    check_images(rimg, gimg)


def test_sixteen_bit(flame0):
    rimg, gimg, *_ = run_case(small(flame0), bytes_per_channel=2)
    assert rimg.dtype == np.uint16
    check_images(rimg, gimg, lsb=256)


def test_rgba_transparency(flame0):
    rimg, gimg, *_ = run_case(small(flame0), nchan=4, transp=1)
    check_images(rimg, gimg)
    rimg, gimg, *_ = run_case(small(flame0), nchan=4, transp=0)
    check_images(rimg, gimg)
    assert (gimg[..., 3] == 255).all()


def test_tone_map_parameters(flame0):
    # vibrancy < 1, highlight power >= 0 (hsv desaturation path), background, gamma threshold
    xml = small(flame0, vibrancy=0.6, highlight_power=1.5, background="0.2 0.1 0.3",
                gamma=2.5, gamma_threshold=0.05, brightness=12)
    rimg, gimg, *_ = run_case(xml)
    check_images(rimg, gimg)


@pytest.mark.parametrize("shape", ["hermite", "box", "triangle", "bell", "bspline", "mitchell",
                                   "blackman", "catrom", "hanning", "hamming", "lanczos3",
                                   "lanczos2", "quadratic"])
def test_spatial_filter_shapes(flame0, shape):
    rimg, gimg, *_ = run_case(small(flame0, filter_shape=shape, filter=0.8))
    check_images(rimg, gimg)


def test_interlaced_fields(flame0):
    xml = small(flame0)
    for field in (1, 2):
        rimg, gimg, *_ = run_case(xml, field=field)
        check_images(rimg, gimg)


def test_pixel_aspect(flame0):
    rimg, gimg, *_ = run_case(small(flame0), pixel_aspect_ratio=1.2)
    check_images(rimg, gimg)


@pytest.mark.parametrize("variant", ["os1", "os2_earlyclip"])
def test_row_strip_postpass_equals_whole_frame(flame0, variant):
    """Multi-GPU post-pass: every rank runs density estimation, filtering and
    tone mapping for its strip of image rows only. The strips of all "ranks" (run here one
    after the other on one GPU, from the same histogram) must tile the whole-frame image
    byte for byte."""
    from flam3_b200.multigpu import strip_rows
    ref, gpu, core = libs()
    if variant == "os1":
        xml, kw = resize(flame0, 200, 150, scale=75, quality=20), {}
    else:
        xml, kw = resize(flame0, 160, 122, scale=60, quality=20, oversample=2), {"earlyclip": 1}
    gr, gg = ref.parse(xml), gpu.parse(xml)
    _, _, buckets, _, geo, _ = ref_render_capture(ref, gr, **kw)
    whole, _ = gpu_render_from_histogram(gpu, core, gg, buckets, geo, **kw)
    H = whole.shape[0]
    nranks = 3
    tiled = np.zeros_like(whole)
    calls = []
    try:
        gpu.set_exchange(lambda ctx: None)
        gpu.set_strip_gather(lambda ctx, image, row_bytes, rows_per: calls.append((row_bytes, rows_per)))
        for r in range(nranks):
            gpu.set_shard(r, nranks)
            img, _ = gpu_render_from_histogram(gpu, core, gg, buckets, geo, **kw)
            r0, r1, per = strip_rows(H, r, nranks)
            tiled[r0:r1] = img[r0:r1]
    finally:
        gpu.set_shard(0, 1)
        gpu.set_exchange(None)
        gpu.set_strip_gather(None)
    assert len(calls) == nranks and calls[0] == (whole.shape[1] * whole.shape[2], (H + 2) // 3)
    assert np.array_equal(tiled, whole), "%d bytes differ" % int((tiled != whole).sum())


@pytest.mark.parametrize("variant", ["default", "tone", "rgba", "earlyclip", "os2"])
def test_float_first_final_pass_gives_the_double_bytes(flame0, variant):
    """The final pass computes every pixel in float and redoes in double the few % whose
    channels land within 0.01 of a truncation boundary (f3b_post.cu). The image must be the
    one the all-double pass (F3B_FINAL_DOUBLE=1) produces, byte for byte."""
    import os
    _, gpu, _ = libs()
    kw = {}
    if variant == "tone":
        xml = small(flame0, quality=100, vibrancy=0.6, highlight_power=1.5, background="0.2 0.1 0.3",
                    gamma=2.5, gamma_threshold=0.05, brightness=12)
    elif variant == "os2":
        xml = resize(flame0, 640, 480, scale=240, quality=100, oversample=2, temporal_samples=1)
    else:
        xml = resize(flame0, 640, 480, scale=240, quality=100, temporal_samples=1)
    if variant == "rgba":
        kw = dict(nchan=4, transp=1)
    if variant == "earlyclip":
        kw = dict(earlyclip=1)
    gg = gpu.parse(xml)
    try:
        os.environ["F3B_FINAL_DOUBLE"] = "1"
        a, sa, _ = gpu.render(gg, isaac_seed="ff", **kw)
    finally:
        os.environ.pop("F3B_FINAL_DOUBLE", None)
    b, sb, _ = gpu.render(gg, isaac_seed="ff", **kw)
    assert sa["rv"] == 0 and sb["rv"] == 0 and a.mean() > 1.0
    assert np.array_equal(a, b), "%d bytes differ" % int((a != b).sum())
