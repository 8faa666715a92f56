"""Integer bucket semantics (`bucket` types, `bump_no_overflow`).

The reference's bits 32/33 buckets are uint32: every add is truncated, sums are exact up to
UINT_MAX and saturate there (flam3.c:3838-3854). The GPU adds into float staging planes (one
vector reduction per hit), which are exact only below 2^24, and sweeper warps move whatever
grows large into uint32 buckets while the walkers run (f3b_iterate.cu). These tests pin:

 * carries change nothing: the histogram with sweepers carrying at a tiny threshold equals,
   bucket for bucket, the one rendered without sweepers;
 * sums far above 2^24 stay exact: regrouping the adds (warp aggregation on / off, three
   shards against one) gives identical buckets, which rounding float sums would not;
 * saturation at UINT_MAX, through the merge and through the sweepers' wrapped atomics;
 * bits=64: double buckets, adds not truncated.
"""
import os

import numpy as np
import pytest

from flam3_b200.genome_xml import resize, set_flame_attrs
from parity_util import libs

pytestmark = pytest.mark.gpu

COLLAPSED = """<flame time="0" size="128 128" center="0 0" scale="32" quality="%g" temporal_samples="1"
   estimator_radius="0">
   <xform weight="1" color="0" linear="1" coefs="0.2 0 0 0.2 0.4 0.1"/>
   <xform weight="1" color="1" linear="1" coefs="0.1 0.05 -0.05 0.1 0.38 0.12"/>
   <color index="0" rgb="255 128 64"/><color index="255" rgb="32 64 255"/></flame>"""


def _hist(gpu, ctx, gg, env, **kw):
    old = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            os.environ[k] = v
        _, st, _ = gpu.render(gg, **kw)
        assert st["rv"] == 0
        return ctx.hist_download_u32().astype(np.int64), ctx.exactness()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _ctx(gpu, core, gg, **kw):
    geo = gpu.buffer_geometry(gg, **kw)
    ctx = gpu.render_context(core)
    ctx.width, ctx.height = geo["width"], geo["height"]
    return ctx


@pytest.mark.parametrize("variant", ["alpha_is_count", "count_plane"])
def test_carries_leave_the_histogram_unchanged(flame0, variant):
    _, gpu, core = libs()
    xml = resize(flame0, 160, 120, scale=60, quality=200, temporal_samples=1)
    if variant == "count_plane":
        xml = set_flame_attrs(xml, palette_mode="linear")
    gg = gpu.parse(xml)
    ctx = _ctx(gpu, core, gg)
    plain, _ = _hist(gpu, ctx, gg, {"F3B_SWEEP": "0"}, isaac_seed="carry")
    swept, (lost, sat) = _hist(gpu, ctx, gg, {"F3B_SWEEP": "1", "F3B_CARRY": "600"}, isaac_seed="carry")
    assert plain.sum() > 0 and plain.max() > 600 * 50   # many carries per bucket happened
    assert lost == 0 and sat == 0
    assert np.array_equal(plain, swept), int((plain != swept).any(axis=-1).sum())


def test_sums_above_2_pow_24_are_exact():
    """A collapsed attractor at a quality that puts ~1e8 into single uint32 channels. Such a
    bucket fills faster than a staging copy can be emptied, so the library must pick the
    aggregating kernels by itself (probe), whose multi-lane sums go straight into the integer
    buckets. Regrouping the adds - three shards against one - gives identical buckets (rounded
    float sums would not), and the bookkeeping reports no staged value at 2^24."""
    _, gpu, core = libs()
    gg = gpu.parse(COLLAPSED % 400.0)
    ctx = _ctx(gpu, core, gg)
    a, (lost_a, _) = _hist(gpu, ctx, gg, {}, isaac_seed="big", sub_batch_size=500)
    b, (lost_b, _) = _hist(gpu, ctx, gg, {"F3B_AGG": "1"}, isaac_seed="big", sub_batch_size=500)
    assert a.max() > 4 * 2 ** 24, a.max()
    assert lost_a == 0 and lost_b == 0
    assert np.array_equal(a, b), int((a != b).any(axis=-1).sum())
    # every add is a whole palette entry: the count channel is a multiple of 255
    assert (a[..., 4] % 255 == 0).all()
    parts = []
    try:
        for r in range(3):
            gpu.set_shard(r, 3)
            h, (lost, _) = _hist(gpu, ctx, gg, {}, isaac_seed="big", sub_batch_size=500)
            assert lost == 0
            parts.append(h)
    finally:
        gpu.set_shard(0, 1)
    assert np.array_equal(parts[0] + parts[1] + parts[2], a)


def test_headline_frame_is_exact(flame0):
    """C2 (1080p, quality 2000): the fullest bucket collects 2.1e9 per channel, 125x what a
    float sum holds exactly. No staged value may be seen at 2^24, the count channel must be a
    whole number of hits everywhere (every add is 255), and its total must be the number of
    in-frame iterations."""
    _, gpu, core = libs()
    gg = gpu.parse(resize(flame0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1))
    ctx = _ctx(gpu, core, gg)
    h, (lost, sat) = _hist(gpu, ctx, gg, {}, isaac_seed="c2-exact")
    assert h[..., 4].max() > 50 * 2 ** 24
    assert lost == 0 and sat == 0
    assert (h[..., 4] % 255 == 0).all()
    assert (h[..., 3] == h[..., 4]).all()   # alpha doubles as the count for this genome
    hits = int(h[..., 4].sum() // 255)
    assert 0.74 * 4147200000 < hits < 0.76 * 4147200000


def test_histogram_in_dram_is_exact(flame0):
    """A 4K oversample-2 still: 7718x4358 buckets, 538 MB of staging - the histogram lives in
    DRAM, a cold pass of the sweepers over it takes ~8 ms beside the walkers' scatter, and the
    fullest buckets pass 2^24 in less. What keeps them exact is the walkers' own reports (one in
    512 of the entries they add to, examined by the CTA's sweeper): without them 13-18 staged
    values were seen at 2^24 (profiles/r02_sweeper_big_histograms.txt)."""
    _, gpu, core = libs()
    gg = gpu.parse(resize(flame0, 3840, 2160, scale=1080, quality=500, oversample=2, temporal_samples=1))
    ctx = _ctx(gpu, core, gg)
    h, (lost, sat) = _hist(gpu, ctx, gg, {"F3B_BINNED": "0"}, isaac_seed="dram")
    assert h.shape[:2] == (4358, 7718)
    assert h[..., 4].max() > 4 * 2 ** 24
    assert lost == 0 and sat == 0
    assert (h[..., 4] % 255 == 0).all()
    hits = int(h[..., 4].sum() // 255)
    assert 0.70 * 4147200000 < hits < 0.80 * 4147200000


@pytest.mark.parametrize("through", ["merge", "sweepers"])
def test_uint32_buckets_saturate(flame0, through):
    """Start from buckets 300 below UINT_MAX: wherever the frame adds more than that the
    bucket must end at exactly UINT_MAX (flam3.c:3838-3840), elsewhere at base + adds."""
    _, gpu, core = libs()
    xml = resize(flame0, 96, 72, scale=36, quality=40, temporal_samples=1)
    gg = gpu.parse(xml)
    ctx = _ctx(gpu, core, gg)
    env = {"F3B_SWEEP": "0"} if through == "merge" else {"F3B_SWEEP": "1", "F3B_CARRY": "256"}
    base = np.full((ctx.height, ctx.width, 5), 2 ** 32 - 1 - 300, dtype=np.uint32)
    zeros = np.zeros_like(base)
    try:
        gpu.preload_histogram(zeros)
        adds, _ = _hist(gpu, ctx, gg, {"F3B_SWEEP": "0"}, isaac_seed="sat")
        gpu.preload_histogram(base)
        got, (lost, sat) = _hist(gpu, ctx, gg, env, isaac_seed="sat")
    finally:
        gpu.preload_histogram(None)
    want = np.minimum(base.astype(np.int64) + adds, 2 ** 32 - 1)
    assert (adds > 300).any() and (adds == 0).any()
    assert lost == 0
    assert np.array_equal(got, want), int((got != want).any(axis=-1).sum())
    if through == "sweepers":
        assert sat > 0


def test_bits_64_double_buckets(flame0):
    """bits=64: adds are not truncated (bucket_double, flam3.c:3724). With a vis < 1 xform the
    per-hit values are fractional: the integer buckets of bits=33 truncate each add, the
    double buckets keep the fractions, so their totals are larger by about half a unit per
    channel and hit; carrying through the sweepers changes the doubles only by rounding."""
    _, gpu, core = libs()
    xml = resize(flame0, 160, 120, scale=60, quality=100, temporal_samples=1)
    xml = xml.replace("<xform ", '<xform opacity="0.6" ', 1)
    gg = gpu.parse(xml)
    ctx = _ctx(gpu, core, gg)
    h33, _ = _hist(gpu, ctx, gg, {}, isaac_seed="d", bits=33)
    h64, _ = _hist(gpu, ctx, gg, {"F3B_SWEEP": "0"}, isaac_seed="d", bits=64)
    h64c, _ = _hist(gpu, ctx, gg, {"F3B_SWEEP": "1", "F3B_CARRY": "900"}, isaac_seed="d", bits=64)
    assert h33.sum() > 0
    s33, s64, s64c = (h[..., :4].sum(dtype=np.float64) for h in (h33, h64, h64c))
    assert s64 > s33 * 1.0005          # truncation removed a visible amount
    assert abs(s64 - s64c) <= 2e-4 * s64  # downloads truncate per bucket; sums agree


@pytest.mark.parametrize("variant", ["alpha_is_count", "count_plane", "opacity_xaos_final", "temporal_samples",
                                     "one_region"])
def test_binned_scatter_gives_the_same_histogram(flame0, flame1, variant):
    """Histograms larger than L2 are filled through per-region record queues and an
    L2-resident tile (f3b_iterate.cu, binned scatter). Forced here on small frames, with
    regions of 0.25 MB and queues so small that a frame needs many launches: the buckets must
    equal those of the direct path bit for bit, also for one shard of three."""
    _, gpu, core = libs()
    kw = {}
    xml = resize(flame0, 320, 240, scale=120, quality=60, temporal_samples=1)
    if variant == "count_plane":
        xml = set_flame_attrs(xml, palette_mode="linear")
    elif variant == "opacity_xaos_final":
        xml = xml.replace("<xform ", '<xform opacity="0.6" chaos="1 0.3 1" ', 1)
        xml = xml.replace("</flame>", '<finalxform color="0" symmetry="1" linear="0.9" spherical="0.1" '
                          'coefs="1 0 0 1 0.05 0"/></flame>')
    elif variant == "temporal_samples":
        xml = "<flames>" + resize(flame0, 320, 240, scale=120, quality=60, temporal_samples=700) + \
              resize(flame1, 320, 240, scale=120, quality=60, temporal_samples=700).replace('time="0"', 'time="100"', 1) + "</flames>"
        kw = {"ngenomes": 2, "time": 50.0}
    gg = gpu.parse(xml)
    ctx = _ctx(gpu, core, gg, **kw)
    binned_env = {"F3B_BINNED": "1", "F3B_BIN_REGION_MB": "64" if variant == "one_region" else "0.25",
                  "F3B_BIN_MEM_MB": "160"}
    direct, _ = _hist(gpu, ctx, gg, {"F3B_BINNED": "0"}, isaac_seed="bin", sub_batch_size=400, **kw)
    binned, (lost, _) = _hist(gpu, ctx, gg, binned_env, isaac_seed="bin", sub_batch_size=400, **kw)
    assert direct.sum() > 0 and lost == 0
    assert np.array_equal(direct, binned), int((direct != binned).any(axis=-1).sum())
    try:
        gpu.set_shard(1, 3)
        d1, _ = _hist(gpu, ctx, gg, {"F3B_BINNED": "0"}, isaac_seed="bin", sub_batch_size=400, **kw)
        b1, _ = _hist(gpu, ctx, gg, binned_env, isaac_seed="bin", sub_batch_size=400, **kw)
    finally:
        gpu.set_shard(0, 1)
    assert 0 < d1.sum() < direct.sum()
    assert np.array_equal(d1, b1)
