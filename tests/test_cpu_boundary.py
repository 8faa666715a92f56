"""CPU tests of the drop-in boundary and the host logic: the C-ABI libraries load and export
every symbol their headers declare, struct sizes are the reference's, error paths that do
not need a GPU behave like the reference, and the multi-GPU sharding rule partitions the
work (world_size 2, gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(f3b_\w+|flam3_render\w*)\s*\(", txt)))


def test_core_library_exports_every_declared_symbol(core):
    names = declared_functions("flam3_b200.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(core._dll, n), "libflam3_b200.so lacks %s" % n
    from flam3_b200.capi import CORE_SYMBOLS

    assert set(CORE_SYMBOLS) <= set(names)
    assert core._dll.f3b_abi_version() == 3


def test_dropin_exports_the_reference_entry_points(gpu):
    for n in declared_functions("flam3_b200_dropin.h"):
        assert hasattr(gpu._dll, n), "drop-in libflam3.so lacks %s" % n
    # and the rest of flam3.h still resolves in the same library (parser, interpolation ...)
    for n in ("flam3_parse_xml2", "flam3_interpolate", "flam3_init_frame", "flam3_print_to_string",
              "flam3_iterate", "flam3_make_strip", "flam3_estimate_bounding_box"):
        assert hasattr(gpu._dll, n)


def test_frontend_symbols_are_exported(gpu):
    """include/flam3_b200_frontend.h: every function it declares is in the drop-in library (and
    linked into f3b-render)."""
    names = declared_functions("flam3_b200_frontend.h")
    assert {"f3b_parse_flam3", "f3b_get_palette", "f3b_interpolate", "f3b_print_flam3", "f3b_write_ppm",
            "f3b_genome_copy"} <= set(names)
    for n in names:
        assert hasattr(gpu._dll, n), "drop-in libflam3.so lacks %s" % n
    out = subprocess.run(["nm", os.path.join(ROOT, "dropin", "_build", "f3b-render")], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(r" T %s$" % n, out, flags=re.M), "f3b-render lacks %s" % n


def test_struct_sizes_are_the_reference_abi(gpu):
    # layouts are ABI
    assert gpu.abi_sizes() == dict(randctx=288, flam3_xform=3240, flam3_genome=10728,
                                   flam3_frame=360)


def test_memory_required_matches_reference_formula(gpu, flame0):
    g = gpu.parse(flame0)
    # os^2 * W * H * bytes * 9 (flam3.c:3884-3886), bits 33 counts as 32
    assert gpu.memory_required(g, 0, 33) == 640 * 480 * 4 * 9.0
    assert gpu.memory_required(g, 0, 64) == 640 * 480 * 8 * 9.0


def test_memory_required_matches_reference(gpu, ref, flame0):
    from flam3_b200.genome_xml import resize

    xml = resize(flame0, 800, 600, oversample=3)
    for bits in (32, 33, 64):
        assert gpu.memory_required(gpu.parse(xml), 0, bits) == ref.memory_required(ref.parse(xml), 0, bits)


def test_bad_bits_is_rejected_before_any_gpu_work(gpu, flame0):
    g = gpu.parse(flame0)
    img, st, _ = gpu.render(g, bits=16)
    assert st["rv"] == 1 and img.max() == 0


def test_no_cpu_fallback_without_gpu(gpu, flame0):
    """On a box without a usable GPU the render must fail loudly (return 1), never fall
    back to a CPU path. (On the GPU box this test is skipped.)"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = gpu.parse(flame0)
    img, st, _ = gpu.render(g)
    assert st["rv"] == 1 and img.max() == 0


def test_product_does_not_link_the_oracle():
    out = subprocess.run(["ldd", os.path.join(ROOT, "dropin", "_build", "libflam3.so")],
                         capture_output=True, text=True).stdout
    assert "oracle" not in out and "libflam3_ref" not in out
    out = subprocess.run(["ldd", os.path.join(ROOT, "flam3_b200", "lib", "libflam3_b200.so")],
                         capture_output=True, text=True).stdout
    assert "oracle" not in out and "flam3_ref" not in out


def test_packed_sample_layout(gpu, flame0):
    """Host packer: header fields, 16-byte aligned xform records, pre-truncated palette."""
    f = gpu._dll.f3b_debug_pack_genome
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    g = gpu.parse(flame0)
    w = np.zeros(12288, dtype=np.uint32)
    used = C.c_size_t()
    assert f(g.ptr, 0, 1, w.ctypes.data, w.size, C.byref(used)) == 0
    fl = w.view(np.float32)
    assert w[1] == 4 and w[2] == 0xFFFFFFFF and w[3] == 0   # 4 xforms, no final, no xaos
    assert w[23] == 4 and w[24] == 3 and w[25] == 4          # nxf, plot-ready | packed, chooser stride
    xftab = w[w[21]: w[21] + 4]
    assert all(int(o) % 4 == 0 for o in xftab)
    # xform 0: spherical only -> fixed-function slot, empty variation list
    rec = int(xftab[0])
    assert (w[rec + 9] >> 8) == 0 and fl[rec + 10] == 0.0 and fl[rec + 11] == 1.0
    assert abs(fl[rec + 0] - (-0.681206)) < 1e-6 and abs(fl[rec + 3] - (-0.0779465)) < 1e-6
    pal = fl[w[22]: w[22] + 1024].reshape(256, 4)
    assert np.all(pal == np.floor(pal)) and np.all(pal[:, 3] == 255.0)
    # the rgba8 copy behind it (F3B_PLOT_PACKED): the same whole numbers, a byte each
    rgba = w[w[22] + 1024: w[22] + 1280]
    assert int(w[0]) == int(w[22]) + 1280
    for k in range(4):
        assert np.array_equal((rgba >> (8 * k)) & 0xFF, pal[:, k].astype(np.uint32))
    # the chooser row: four equal weights -> 4096 / 8192 / 12288 / 16384
    assert list(w[w[20]: w[20] + 4]) == [4096, 8192, 12288, 16384]


@pytest.mark.parametrize("cold", [False, True])
@pytest.mark.parametrize("mode", ["hsv_circular", "hsv", "rgb", "sweep", "blend"])
def test_cached_interpolation_equals_flam3_interpolate(gpu, flame0, flame1, mode, cold, monkeypatch):
    """The genome of a temporal sample (f3b_render.c, f3b_interpolate_cached: aligned copies and
    the palettes' HSV values kept across the samples of a frame) is byte for byte what
    flam3_interpolate makes for the same time: genome struct, xform arrays, chaos matrix -
    inside a segment, on and beyond the key frames, with a final xform on one side only
    (padding) and for every palette interpolation mode."""
    from types import SimpleNamespace
    from flam3_b200.genome_xml import set_flame_attrs
    from test_cpu_frontend import _raw
    # warm (default): the call under test follows another one on the same cache and result, as
    # one temporal sample follows another in a packer thread (the result is reset in place);
    # cold: a single call from scratch
    if cold:
        monkeypatch.setenv("F3B_DEBUG_INTERP_COLD", "1")
    a = flame0.replace("</flame>", '<finalxform color="0" symmetry="1" linear="0.9" spherical="0.1" '
                       'coefs="1 0 0 1 0.05 0"/></flame>')
    b = flame1.replace('time="0"', 'time="100"', 1)
    extra = {"hsv_circular": {}, "hsv": {"palette_interpolation": "hsv"}, "rgb": {"palette_interpolation": "rgb"},
             "sweep": {"palette_interpolation": "sweep"},
             "blend": {"palette_interpolation": "hsv_circular", "hsv_rgb_palette_blend": "0.35"}}[mode]
    if extra:
        a, b = set_flame_attrs(a, **extra), set_flame_attrs(b, **extra)
    c = set_flame_attrs(flame0, **extra).replace('time="0"', 'time="250"', 1) if extra else flame0.replace('time="0"', 'time="250"', 1)
    g = gpu.parse("<flames>" + a + b + c + "</flames>")
    assert g.n == 3
    d = gpu._dll
    size = gpu.abi_sizes()["flam3_genome"]
    d.flam3_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    d.flam3_interpolate.restype = None
    d.f3b_debug_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    d.clear_cp.argtypes = [C.c_void_p, C.c_int]
    for t in (-3.0, 0.0, 0.001, 37.25, 50.0, 99.999, 100.0, 100.5, 180.0, 250.0, 260.0):
        ra, rb = (C.c_char * size)(), (C.c_char * size)()
        d.flam3_interpolate(g.ptr, 3, t, 0.0, ra)
        assert d.f3b_debug_interpolate(g.ptr, 3, t, rb) == 0
        (ga, xa, ca), = _raw(gpu, SimpleNamespace(ptr=C.addressof(ra), n=1))
        (gb, xb, cb), = _raw(gpu, SimpleNamespace(ptr=C.addressof(rb), n=1))
        diff = [i for i in range(len(ga)) if ga[i] != gb[i]]
        assert not diff, "t=%g: genome bytes differ at %s" % (t, diff[:12])
        assert ca == cb and np.array_equal(xa, xb), t
        d.clear_cp(ra, 0)
        d.clear_cp(rb, 0)


@pytest.mark.parametrize("grid", [(1938, 1098), (658, 498), (1042, 1042), (97, 13)])
def test_staging_layout_is_a_bijection(grid):
    """Inside a staging copy bucket b sits at (b & ~m) | ((b * a) & m): a
    bijection of every aligned block of m + 1 entries, undone with the inverse multiplier, the
    copy a whole number of blocks, neighbouring buckets at least 128 bytes (8 entries) apart."""
    import flam3_b200 as fb
    core = fb.load_core()
    out = (C.c_uint64 * 5)()
    core._dll.f3b_debug_stage_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    assert core._dll.f3b_debug_stage_layout(grid[0], grid[1], 0, out) == 0
    stride, copies, m, a, inv = (int(v) for v in out)
    assert stride >= grid[0] * grid[1] and 1 <= copies <= 4
    assert m == 255 and stride % (m + 1) == 0
    assert (a * inv) % 2 ** 32 == 1
    b = np.arange(0, 4 * (m + 1), dtype=np.uint64) + 7 * (m + 1)
    s = (b & ~np.uint64(m)) | ((b * np.uint64(a)) & np.uint64(m))
    assert len(set(s.tolist())) == len(b) and (s >> 8 == b >> 8).all()
    back = (s & ~np.uint64(m)) | (((s * np.uint64(inv)) & np.uint64(0xFFFFFFFF)) & np.uint64(m))
    assert (back == b).all()
    gap = np.abs(s[1:].astype(np.int64) - s[:-1].astype(np.int64))[(b[1:] >> 8) == (b[:-1] >> 8)]
    assert gap.min() >= 8


def test_shard_rule_partitions_groups():
    from flam3_b200.multigpu import shard_groups

    for n_groups in (1, 7, 592, 1777):
        for nranks in (1, 2, 3, 8):
            parts = [shard_groups(n_groups, r, nranks) for r in range(nranks)]
            allg = sorted(g for p in parts for g in p)
            assert allg == list(range(n_groups))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r})
from flam3_b200.multigpu import shard_groups
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ["MASTER_ADDR"] = "127.0.0.1"
os.environ["MASTER_PORT"] = sys.argv[3]
dist.init_process_group("gloo", rank=rank, world_size=world)
# every chain group g deposits (g+1) hits into bucket g % nb: a stand-in for the kernel that
# makes the exchanged histogram predictable
n_groups, nb = 101, 16
hist = np.zeros(nb * 4, dtype=np.float32)
for g in shard_groups(n_groups, rank, world):
    hist[(g % nb) * 4: (g % nb) * 4 + 4] += float(g + 1)
t = torch.from_numpy(hist)
dist.all_reduce(t, op=dist.ReduceOp.SUM)
want = np.zeros(nb * 4, dtype=np.float32)
for g in range(n_groups):
    want[(g % nb) * 4: (g % nb) * 4 + 4] += float(g + 1)
assert np.array_equal(t.numpy(), want), (rank, t.numpy()[:8], want[:8])
dist.destroy_process_group()
print("ok", rank)
"""


def test_histogram_exchange_world_size_2_gloo(tmp_path):
    """The N>1 path: private histograms of disjoint chain-group shards, summed by one
    all-reduce, equal the single-process histogram. gloo on CPU, world_size 2."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT))
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", port],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


@pytest.mark.parametrize("height,nranks", [(1080, 8), (480, 3), (5, 8), (1, 2), (7, 7)])
def test_strip_rows_tile_the_image(height, nranks):
    """Row strips of the multi-GPU post-pass: equal blocks of ceil(h/n) rows, clipped, that
    cover every image row exactly once (the rule f3b_render.c applies per rank)."""
    from flam3_b200.multigpu import strip_rows
    seen = []
    for r in range(nranks):
        r0, r1, per = strip_rows(height, r, nranks)
        assert per == -(-height // nranks) and 0 <= r0 <= r1 <= height and r1 - r0 <= per
        seen += list(range(r0, r1))
    assert seen == list(range(height))


@pytest.mark.parametrize("first,last,dtime,n", [(0, 99, 1, 8), (10, 12, 1, 8), (0, 20, 3, 4), (5, 5, 1, 2)])
def test_animate_frame_plan_covers_every_frame_once(first, last, dtime, n):
    """flam3_b200.animate: worker r renders first + r*dtime, then every n*dtime — together
    exactly flam3-animate's own frame list (flam3-animate.c:225)."""
    from flam3_b200.animate import frames_of, plan_frames
    per_worker = frames_of(plan_frames(first, last, dtime, n), last)
    flat = sorted(t for w in per_worker for t in w)
    assert flat == list(range(first, last + 1, dtime))
    assert max(len(w) for w in per_worker) - min(len(w) for w in per_worker) <= 1


def test_animate_frame_range_follows_flam3_animate():
    from flam3_b200.animate import frame_range
    xml = '<flames><flame time="0" size="1 1"/><flame time="100" size="1 1"/></flames>'
    assert frame_range({}, xml) == (0, 99)
    assert frame_range({"begin": "5"}, xml) == (5, 99)
    assert frame_range({"begin": "5", "end": "7"}, xml) == (5, 7)
    assert frame_range({"frame": "42"}, xml) == (42, 42)
    assert frame_range({"time": "17"}, xml) == (17, 17)


def test_chooser_closed_form_equals_the_reference_loop():
    """The packer finds where the reference's 16384-step running sum (flam3.c:197-218,
    r += dr in double) crosses the cumulative weights in closed form, binade by binade
    (f3b_pack.c: first_crossings). It must give the very indices the plain loop gives:
    random weights, weights with short mantissas (round-to-even ties), tiny and zero
    weights, many xforms."""
    import flam3_b200 as fb
    gpu = fb.load_dropin()
    f = gpu._dll.f3b_debug_chooser_counts
    f.argtypes = [C.c_double, C.c_uint, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_uint32), C.c_int]
    rs = np.random.RandomState(2024)
    cases = []
    for _ in range(1500):
        n = rs.randint(2, 9)
        kind = rs.randint(6)
        if kind == 0:
            w = rs.uniform(0.0, 1.0, n)
        elif kind == 1:
            w = rs.randint(1, 9, n) / 8.0                    # short mantissas: ties
        elif kind == 2:
            w = 2.0 ** rs.randint(-12, 4, n)                 # powers of two
        elif kind == 3:
            w = rs.uniform(0.0, 1.0, n) * (rs.uniform(0, 1, n) < 0.7)   # some zero weights
            w[0] = max(w[0], 1e-3)
        elif kind == 4:
            w = rs.uniform(0.0, 1.0, n) * 10.0 ** rs.randint(-8, 3, n)  # wildly different scales
        else:
            w = rs.randint(1, 4, n) * 3.0 / 16384.0          # steps of a few grains
        cases.append(w)
    cases.append(np.full(40, 0.37))
    cases.append(np.array([1.0, 1e-12, 1.0]))
    cases.append(np.array([0.5, 0.5]))
    for w in cases:
        total = 0.0
        for d in w:
            total += float(d)
        dr = total / 16384.0
        thr, t = [], 0.0
        for d in w[:-1]:
            t += float(d)
            thr.append(t)
        thr_c = (C.c_double * len(thr))(*thr)
        a = (C.c_uint32 * len(thr))()
        b = (C.c_uint32 * len(thr))()
        f(dr, 16384, thr_c, len(thr), a, 1)
        f(dr, 16384, thr_c, len(thr), b, 0)
        assert list(a) == list(b), (w, list(a), list(b))


XAOS_GENOME = """
<flame time="0" size="64 64" center="0 0" scale="16" quality="5">
   <xform weight="0.31" color="0" linear="1" coefs="0.5 0 0 0.5 -0.5 0" chaos="1 0.25 3"/>
   <xform weight="0.07" color="0.5" spherical="1" coefs="0.5 0 0 0.5 0.5 0" chaos="0 1 1"/>
   <xform weight="1.7" color="1" swirl="0.5" linear="0.5" coefs="0.4 0.2 -0.2 0.4 0 0.5" chaos="2 2 0.01"/>
   <finalxform color="0" linear="1" coefs="1 0 0 1 0 0"/>
   <color index="0" rgb="255 0 0"/>
</flame>
"""


def _chooser_from_pack(gpu, genomes, idx=0):
    """Expands the packed cumulative counts of a sample back into the 16384-entry rows."""
    f = gpu._dll.f3b_debug_pack_genome
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    w = np.zeros(12288, dtype=np.uint32)
    used = C.c_size_t()
    assert f(genomes.ptr, idx, 1, w.ctypes.data, w.size, C.byref(used)) == 0
    nstd, xaos, stride, off = int(w[1]), int(w[3]), int(w[25]), int(w[20])
    rows = nstd + 1 if xaos else 1
    out = np.empty((rows, 16384), dtype=np.uint16)
    k = np.arange(16384)
    for r in range(rows):
        cum = w[off + r * stride: off + r * stride + nstd].astype(np.int64)
        assert np.all(np.diff(cum) >= 0) and cum[-1] == 16384
        assert np.all(w[off + r * stride + nstd: off + (r + 1) * stride] == 16384)   # padding
        out[r] = np.searchsorted(cum, k, side="right")   # number of thresholds <= k
    return out


def _reference_chooser(lib, genomes, idx=0):
    f = lib._dll.h_xform_distrib
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    buf = np.zeros(64 * 16384, dtype=np.uint16)
    rows = C.c_int()
    assert f(genomes.ptr, idx, buf.ctypes.data, C.byref(rows)) == 0
    return buf[: rows.value * 16384].reshape(rows.value, 16384)


def test_chooser_tables_equal_the_reference_entry_for_entry(gpu, ref, flame0):
    """flam3_create_xform_distrib / flam3_create_chaos_distrib (flam3.c:76-224) against the
    packed cumulative counts the kernel compares its 14-bit draw with: every one of the 16384
    entries of every row, for test.flam3, an xaos + final-xform genome and all 122 random
    genomes of config C5."""
    import glob
    cases = [flame0, XAOS_GENOME]
    cases += [open(p).read() for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "c5", "c5_*.flam3")))]
    assert len(cases) == 124
    n_xaos = 0
    for xml in cases:
        gr, gg = ref.parse(xml), gpu.parse(xml)
        want = _reference_chooser(ref, gr)
        got = _chooser_from_pack(gpu, gg)
        n_xaos += want.shape[0] > 1
        assert got.shape == want.shape
        assert np.array_equal(got, want), int((got != want).sum())
    assert n_xaos >= 1
