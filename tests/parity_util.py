"""Shared helpers of the parity tests (also used by __graft_entry__.smoke)."""
import ctypes as C
import os

import numpy as np

import flam3_b200 as fb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so")

_cache = {}


def libs():
    if "ref" not in _cache:
        _cache["ref"] = fb.Flam3Lib(REF_PATH)
        _cache["gpu"] = fb.load_dropin()
        _cache["core"] = fb.load_core()
    return _cache["ref"], _cache["gpu"], _cache["core"]


def psnr(a, b, peak=255.0):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else float(10 * np.log10(peak ** 2 / mse))


def ref_render_capture(ref, genomes, **kw):
    """Reference render (nthreads=1 -> deterministic DE) that also returns the
    final histogram and accumulator captured by oracle/ref_wrap.c.

    Returns (image, stats, buckets uint32[h,w,5], accum float32[h,w,4], geometry)."""
    kw.setdefault("nthreads", 1)
    bits = kw.get("bits", 33)
    assert bits == 33, "capture layout is the bits=33 one (uint32[5] buckets, float[4] accum)"
    geo = ref.buffer_geometry(genomes, **{k: v for k, v in kw.items()
                                            if k in ("field", "time", "pixel_aspect_ratio",
                                                     "first_genome", "ngenomes")})
    d = ref._dll
    d.f3ref_capture.argtypes = [C.c_int]
    d.f3ref_captured_block.restype = C.c_void_p
    d.f3ref_captured_block.argtypes = [C.POINTER(C.c_size_t)]
    d.f3ref_capture(1)
    img, st, rc = ref.render(genomes, **kw)
    size = C.c_size_t(0)
    p = d.f3ref_captured_block(C.byref(size))
    n = geo["width"] * geo["height"]
    assert p and size.value >= n * 36, (size.value, n)
    raw = C.string_at(p, n * 36)
    d.f3ref_capture(0)
    buckets = np.frombuffer(raw[: n * 20], dtype=np.uint32).reshape(geo["height"], geo["width"], 5).copy()
    accum = np.frombuffer(raw[n * 20: n * 36], dtype=np.float32).reshape(geo["height"], geo["width"], 4).copy()
    return img, st, buckets, accum, geo, rc


def gpu_render_from_histogram(gpu, core, genomes, buckets, geo, **kw):
    """Drop-in flam3_render with the chaos game replaced by `buckets`; also returns the
    accumulator left on the device."""
    gpu.inject_histogram(buckets)
    try:
        img, st, _ = gpu.render(genomes, **kw)
    finally:
        gpu.inject_histogram(None)
    assert st["rv"] == 0, st
    ctx = gpu.render_context(core)
    ctx.width, ctx.height = geo["width"], geo["height"]
    acc = ctx.accum_download()
    return img, acc


def postpass_parity(xml, **kw):
    """max |GPU - reference| over the output bytes, and the fraction of bytes that differ,
    when both post-passes start from the reference's own histogram."""
    ref, gpu, core = libs()
    gr, gg = ref.parse(xml), gpu.parse(xml)
    rimg, rst, buckets, racc, geo, _ = ref_render_capture(ref, gr, **kw)
    gimg, gacc = gpu_render_from_histogram(gpu, core, gg, buckets, geo, **kw)
    diff = np.abs(gimg.astype(np.int64) - rimg.astype(np.int64))
    return int(diff.max()), float((diff > 0).mean())
