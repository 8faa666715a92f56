"""Generate tests/golden/oracle_golden.npz by RUNNING THE UNMODIFIED REFERENCE here
(oracle/_ref/libflam3_ref.so, compiled from /root/reference). The reference ships no golden
vectors of its own, so these are the known answers the C oracle (and through it the CUDA
path) is pinned to on boxes where the reference tree is absent.

    python tools/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import flam3_b200 as fb  # noqa: E402
from flam3_b200.genome_xml import resize  # noqa: E402
from parity_util import ref_render_capture  # noqa: E402
from variation_cases import VAR_NAMES, genome_xml, test_points  # noqa: E402

XAOS_XML = """
<flame time="0" size="64 64" center="0 0" scale="16" quality="5" temporal_samples="1">
   <xform weight="0.5" color="0" linear="1" coefs="0.5 0 0 0.5 -0.5 0" chaos="1 0.2 1"/>
   <xform weight="0.3" color="0.5" linear="1" coefs="0.5 0 0 0.5 0.5 0" chaos="0.5 1 2"/>
   <xform weight="0.2" color="1" linear="1" coefs="0.5 0 0 0.5 0 0.5" chaos="1 1 0.3"/>
</flame>
"""


class DEH(C.Structure):
    _fields_ = [("max_filtered_counts", C.c_int), ("max_filter_index", C.c_int),
                ("kernel_size", C.c_int), ("filter_widths", C.POINTER(C.c_double)),
                ("filter_coefs", C.POINTER(C.c_double))]


def main():
    ref = fb.Flam3Lib(os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so"))
    d = ref._dll
    out = {}

    # ISAAC stream (isaac.c, LP64 behaviour)
    draws = (C.c_ulong * 200)()
    d.h_isaac_draws(b"flam3b200", 200, draws, None, None)
    out["isaac_seed"] = np.frombuffer(b"flam3b200", dtype=np.uint8)
    out["isaac_draws"] = np.array(list(draws), dtype=np.uint64)

    # chooser tables
    flame0 = open(os.path.join(ROOT, "tests", "golden", "test_flame0.flam3")).read()
    d.h_xform_distrib.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    for key, xml in (("chooser_flame0", flame0), ("chooser_xaos", XAOS_XML)):
        g = ref.parse(xml)
        buf = np.zeros(16384 * 8, dtype=np.uint16)
        rows = C.c_int()
        assert d.h_xform_distrib(g.ptr, 0, buf.ctypes.data, C.byref(rows)) == 0
        out[key] = buf[: 16384 * rows.value].reshape(rows.value, 16384).copy()

    # apply_xform, all 99 variations
    pts = test_points(64)
    d.h_apply_xform.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                C.c_char_p]
    res = np.zeros((99, len(pts), 5))
    for vid in range(99):
        g = ref.parse(genome_xml(vid))
        assert d.h_apply_xform(g.ptr, 0, 0, pts.ctypes.data, len(pts), res[vid].ctypes.data,
                               b"vseed") == 0
    out["apply_points"] = pts
    out["apply_results"] = res

    # a short orbit of flam3_iterate
    d.h_iterate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                            C.c_char_p]
    g = ref.parse(flame0)
    p0 = np.array([0.1, -0.2, 0.5, 1.0])
    samples = np.zeros(4 * 64)
    bad = d.h_iterate(g.ptr, 0, 64, 15, p0.ctypes.data, samples.ctypes.data, b"orbit")
    out["orbit_start"] = p0
    out["orbit_samples"] = samples.reshape(64, 4)
    out["orbit_bad"] = np.array([bad])

    # filter banks
    f = d.flam3_create_de_filters
    f.restype = DEH
    f.argtypes = [C.c_double] * 3 + [C.c_int]
    for key, a in (("de_9_0_04_1", (9.0, 0.0, 0.4, 1)), ("de_5_1_06_3", (5.0, 1.0, 0.6, 3))):
        r = f(*a)
        n = r.max_filter_index + 1
        out[key + "_meta"] = np.array([r.max_filtered_counts, r.max_filter_index, r.kernel_size])
        out[key + "_widths"] = np.ctypeslib.as_array(r.filter_widths, shape=(n,)).copy()
        out[key + "_coefs"] = np.ctypeslib.as_array(r.filter_coefs, shape=(n * r.kernel_size,)).copy()
    sk = d.flam3_spatial_filter
    sk.restype = C.c_double
    sk.argtypes = [C.c_int, C.c_double]
    xs = np.linspace(-3.2, 3.2, 129)
    out["spatial_kernel_x"] = xs
    out["spatial_kernel_y"] = np.array([[sk(k, x) for x in xs] for k in range(14)])
    tf = d.flam3_create_temporal_filter
    tf.restype = C.c_double
    tf.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.POINTER(C.c_double)),
                   C.POINTER(C.POINTER(C.c_double))]
    for t in range(3):
        pw, pd = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        s = tf(12, t, 1.7 if t == 2 else 0.0, 1.2, C.byref(pw), C.byref(pd))
        out["temporal_%d" % t] = np.concatenate([[s], np.ctypeslib.as_array(pw, shape=(12,)),
                                                 np.ctypeslib.as_array(pd, shape=(12,))])

    # one small frame: histogram -> accumulator -> image (nthreads=1, bits=33)
    xml = resize(flame0, 96, 72, scale=36, quality=30, temporal_samples=1)
    g = ref.parse(xml)
    img, st, buckets, accum, geo, _ = ref_render_capture(ref, g, isaac_seed="golden")
    out["frame_xml"] = np.frombuffer(xml.encode(), dtype=np.uint8)
    out["frame_image"] = img
    out["frame_buckets"] = buckets
    out["frame_accum"] = accum
    out["frame_geo"] = np.array([geo[k] for k in ("width", "height", "filter_width", "gutter",
                                                  "de_offset", "oversample")])
    p = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
