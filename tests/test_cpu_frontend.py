"""the repository's own front end (flam3_b200/csrc/f3b_frontend.c) reads
genome files without the reference's parser.c / palettes.c. The oracle is the reference itself:
for every fixture genome, the text the REFERENCE's printer (flam3_print_to_string) produces for
the genome parsed by the own front end must equal, character for character, the text it
produces for the genome parsed by the reference's flam3_parse_xml2. The printer writes every
attribute of the flame, every xform with all its variations and parameters, chaos, final xform
and the 256 palette entries, so equality of the texts is equality of everything a render reads.
No GPU involved."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fixtures():
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.flam3")) +
                   glob.glob(os.path.join(ROOT, "tests", "golden", "c5", "*.flam3")))
    return files


def _printed(lib, genomes):
    return [lib.print_genome(genomes, i) for i in range(genomes.n)]


@pytest.mark.parametrize("path", _fixtures(), ids=lambda p: os.path.basename(p))
def test_own_parser_equals_the_reference_parser(gpu, path):
    xml = open(path).read()
    a = _printed(gpu, gpu.parse(xml))
    b = _printed(gpu, gpu.parse_own(xml))
    assert len(a) == len(b) >= 1
    for x, y in zip(a, b):
        assert x == y


def _raw(lib, genomes):
    """Every data byte a render can read: the genome structs without their pointer fields and
    file name, the xform arrays (variation weights, affine parts, parameters, motion function;
    not what flam3_render derives - precalc values, the active-variation list - nor padding)
    and the chaos matrix."""
    import ctypes as C
    import numpy as np
    d = lib._dll
    ranges = (C.c_int * 16)()
    nr = d.h_genome_nondata_ranges(ranges)
    xranges = (C.c_int * 16)()
    nx = d.h_xform_nondata_ranges(xranges)
    d.h_genome_xforms.restype = C.c_void_p
    d.h_genome_xforms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    d.h_genome_chaos_row.restype = C.POINTER(C.c_double)
    d.h_genome_chaos_row.argtypes = [C.c_void_p, C.c_int, C.c_int]
    sizes = lib.abi_sizes()
    out = []
    for i in range(genomes.n):
        g = bytearray(C.string_at(genomes.ptr + i * sizes["flam3_genome"], sizes["flam3_genome"]))
        for k in range(nr):
            g[ranges[2 * k]:ranges[2 * k] + ranges[2 * k + 1]] = bytes(ranges[2 * k + 1])
        num, fin = C.c_int(), C.c_int()
        xp = d.h_genome_xforms(genomes.ptr, i, C.byref(num), C.byref(fin))
        xs = np.frombuffer(C.string_at(xp, num.value * sizes["flam3_xform"]), dtype=np.uint8).reshape(num.value, -1).copy()
        for k in range(nx):
            xs[:, xranges[2 * k]:xranges[2 * k] + xranges[2 * k + 1]] = 0
        nstd = num.value - (1 if fin.value >= 0 else 0)
        chaos = [[d.h_genome_chaos_row(genomes.ptr, i, r)[c] for c in range(nstd)] for r in range(nstd)]
        # motion elements ride along with the xform array: same masks, one row each
        d.h_xform_motion.restype = C.c_void_p
        d.h_xform_motion.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        rows = [xs]
        for xf in range(num.value):
            nm = C.c_int()
            mp = d.h_xform_motion(genomes.ptr, i, xf, C.byref(nm))
            if nm.value:
                m = np.frombuffer(C.string_at(mp, nm.value * sizes["flam3_xform"]), dtype=np.uint8).reshape(nm.value, -1).copy()
                for k in range(nx):
                    m[:, xranges[2 * k]:xranges[2 * k] + xranges[2 * k + 1]] = 0
                m[:, 0] ^= np.uint8(xf + 1)   # (tie the rows to their xform)
                rows.append(m)
        out.append((bytes(g), np.concatenate(rows), chaos))
    return out


@pytest.mark.parametrize("path", _fixtures(),
                         ids=lambda p: os.path.basename(p))
def test_own_parser_gives_bit_identical_genomes(gpu, path):
    """Stronger than the printed text (which rounds to a few digits): the genome structs, xform
    arrays and chaos matrices of both parsers are the same BYTES, so a render cannot tell them
    apart. This is what catches, e.g., n / 255 against n * (1 / 255) in a palette entry (the
    reference is built with -ffast-math)."""
    import numpy as np
    xml = open(path).read()
    a, b = _raw(gpu, gpu.parse(xml)), _raw(gpu, gpu.parse_own(xml))
    assert len(a) == len(b)
    for (ga, xa, ca), (gb, xb, cb) in zip(a, b):
        diff = [i for i in range(len(ga)) if ga[i] != gb[i]]
        assert not diff, "genome bytes differ at offsets %s" % diff[:20]
        assert ca == cb
        assert xa.shape == xb.shape
        assert np.array_equal(xa, xb), "xform bytes differ at offsets %s" % np.where((xa != xb).any(axis=0))[0][:20]


EXTRA = """
<flames>
<flame name="two  words here" time="3" size="320 200" center="0.25 -0.5" scale="77.5" rotate="200" zoom="0.5"
   supersample="2" filter="0.8" filter_shape="mitchell" temporal_filter_type="gaussian" temporal_filter_width="1.5"
   quality="33" passes="2" temporal_samples="40" background="0.1 0.2 0.3" brightness="5" gamma="3.5"
   highlight_power="1.5" vibrancy="0.7" estimator_radius="7" estimator_minimum="1" estimator_curve="0.6"
   gamma_threshold="0.02" palette_mode="linear" interpolation="smooth" interpolation_type="linear"
   palette_interpolation="sweep" hsv_rgb_palette_blend="0.3" soloxform="1">
   <xform weight="0.5" color="0.3 0.9" symmetry="0.5" coefs="1 0.1 -0.1 1 0.2 0.3" post="0.9 0 0 0.9 0.1 0"
          linear="0.5" julian="0.4" julian_power="3" julian_dist="0.8" chaos="1 2 0.5"/>
   <xform weight="0.3" color="0.6" color_speed="0.2" animate="0" opacity="0.4" var1="13" coefs="0.5 0 0 0.5 0 0"/>
   <xform weight="0.2" color="1" plotmode="off" coefs="0.3 0.3 -0.3 0.3 0.5 0.5" spherical="1" oscilloscope_frequency="2"
          oscilloscope="0.3" Re_A="0.1" mobius="0.2" chaos="0.25"/>
   <finalxform color="0" color_speed="0" coefs="1 0 0 1 0 0" linear="0.8" curl="0.2" curl_c1="0.3"/>
   <color index="0" rgb="255 0 0"/>
   <color index="100" rgba="0 255 0 128"/>
   <color index="200" rgb="0 0 255" a="64"/>
</flame>
<flame time="10" size="64 48" rotate="-100">
   <xform weight="1" color="0" coefs="0.5 0 0 0.5 0 0" linear="1"/>
   <finalxform color="0" coefs="1 0 0 1 0 0" linear="1"/>
   <xform weight="2" color="1" coefs="0.5 0 0 0.5 0.5 0" sinusoidal="1"/>
   <palette count="256" format="RGB">
%s
   </palette>
</flame>
<flame time="20" size="64 48" rotate="170">
   <xform weight="1" color="0" coefs="0.5 0 0 0.5 0 0" linear="1"/>
   <colors count="256" data="%s"/>
</flame>
</flames>
""" % ("\n".join("".join("%02X%02X%02X" % ((i * 8 + j) % 256, (i * 3 + j) % 256, 255 - (i * 8 + j) % 256) for j in range(8))
                 for i in range(32)),
       "".join("00%02x%02x%02x" % (i, 255 - i, (i * 7) % 256) for i in range(256)))


def test_own_parser_on_every_attribute_it_reads(gpu):
    """A hand-written file exercising what the fixtures do not: every <flame> attribute, the
    deprecated spellings (symmetry=, var1=, plotmode=, Re_A=, oscilloscope_*), chaos rows, a
    final xform that is not last, sparse <color> entries (the gaps are interpolated, wrapping
    around), <palette> and <colors> hex blocks, soloxform, smooth interpolation on the first
    segment (reset to linear), rotate values brought within 180 degrees of each other."""
    a = _printed(gpu, gpu.parse(EXTRA))
    b = _printed(gpu, gpu.parse_own(EXTRA))
    assert len(a) == len(b) == 3
    for x, y in zip(a, b):
        assert x == y
    assert 'rotate="260"' in b[1] or re.search(r'rotate="2[56]\d', b[1])
    import numpy as np
    for (ga, xa, ca), (gb, xb, cb) in zip(_raw(gpu, gpu.parse(EXTRA)), _raw(gpu, gpu.parse_own(EXTRA))):
        assert ga == gb and ca == cb and np.array_equal(xa, xb)


PALETTE_LIB_SCRIPT = r"""
import sys
sys.path.insert(0, %(tests)r)
import numpy as np
import flam3_b200 as fb
import test_cpu_frontend as T
lib = fb.load_dropin()
xml = '<flames>' + ''.join(
    '<flame time="%%d" size="32 32" palette="%%d" hue="%%s"><xform weight="1" color="0" coefs="1 0 0 1 0 0" linear="1"/></flame>'
    %% (i, n, h) for i, (n, h) in enumerate([(3, "0"), (3, "0.37"), (11, "-0.2"), (42, "0.999"), (42, "2.5")])) + ''.join(
    '<flame time="%%d" size="32 32"><palette index0="%%d" hue0="%%s" index1="%%d" hue1="%%s" blend="%%s"/>'
    '<xform weight="1" color="0" coefs="1 0 0 1 0 0" linear="1"/></flame>' %% ((10 + i,) + t)
    for i, t in enumerate([(3, "0.1", 42, "0.5", "0.3"), (11, "0", 3, "0", "0.5"), (42, "0.9", 11, "0.2", "0.85")])) + '</flames>'
a, b = T._raw(lib, lib.parse(xml)), T._raw(lib, lib.parse_own(xml))
assert len(a) == len(b) == 8
for (ga, xa, ca), (gb, xb, cb) in zip(a, b):
    diff = [i for i in range(len(ga)) if ga[i] != gb[i]]
    assert not diff, diff[:20]
    assert np.array_equal(xa, xb)
pal = np.frombuffer(a[1][0], dtype=np.uint8)
assert len(set(a[k][0] for k in range(8))) == 8      # the palettes were found, rotated, blended
print("ok")
"""


def test_palette_library_and_hue_rotation(tmp_path):
    """palette="N" hue="h" and the pre-2008 <palette index0= hue0= index1= hue1= blend=/>: the library file (the path comes from the environment variable the
    reference reads, palettes.c:117) and the hexcone rotation, against the reference's parser on
    a library written here. In a fresh process: both sides read the library once."""
    import subprocess
    import sys
    import numpy as np
    rng = np.random.default_rng(5)
    rows = []
    for number in (3, 11, 42):
        cols = rng.integers(0, 256, size=(256, 3))
        cols[:8] = [[0, 0, 0], [255, 255, 255], [7, 7, 7], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [1, 2, 1]]
        data = "\n".join("".join("00%02X%02X%02X" % tuple(c) for c in cols[i:i + 8]) for i in range(0, 256, 8))
        rows.append('<palette number="%d" name="p%d" data="%s"/>' % (number, number, data))
    path = tmp_path / "palettes.xml"
    path.write_text("<palettes>\n" + "\n".join(rows) + "\n</palettes>\n")
    env = dict(os.environ, flam3_palettes=str(path))
    out = subprocess.run([sys.executable, "-c", PALETTE_LIB_SCRIPT % {"tests": os.path.join(ROOT, "tests")}],
                         env=env, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


@pytest.mark.parametrize("path", _fixtures(), ids=lambda p: os.path.basename(p))
def test_own_writer_equals_the_reference_printer(gpu, path):
    """f3b_print_flam3 against flam3_print_to_string on every fixture genome: the same text,
    character for character (attribute order, %g formatting, which variation parameters are
    written, chaos rows cut after their last non-1 entry, rgb / rgba palette lines)."""
    import ctypes as C
    d = gpu._dll
    d.flam3_version.restype = C.c_char_p
    version = b"FLAM3-" + d.flam3_version()
    d.f3b_print_flam3.restype = C.c_void_p
    d.f3b_print_flam3.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    g = gpu.parse(open(path).read())
    size = gpu.abi_sizes()["flam3_genome"]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for i in range(g.n):
        want = gpu.print_genome(g, i)
        ptr = d.f3b_print_flam3(g.ptr + i * size, version, None)
        assert ptr
        got = C.string_at(ptr).decode()
        libc.free(ptr)
        assert got == want


def test_own_writer_on_every_attribute(gpu):
    import ctypes as C
    d = gpu._dll
    d.flam3_version.restype = C.c_char_p
    version = b"FLAM3-" + d.flam3_version()
    d.f3b_print_flam3.restype = C.c_void_p
    d.f3b_print_flam3.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    g = gpu.parse(EXTRA)
    size = gpu.abi_sizes()["flam3_genome"]
    for i in range(g.n):
        got = C.string_at(d.f3b_print_flam3(g.ptr + i * size, version, None)).decode()
        assert got == gpu.print_genome(g, i)
    # and the text reads back as the reference reads its own text (not as the same genome where a
    # palette has alpha: flam3_print writes the premultiplied colours, the parser multiplies again)
    text = "<flames>" + "".join(C.string_at(d.f3b_print_flam3(g.ptr + i * size, version, None)).decode()
                                for i in range(g.n)) + "</flames>"
    mine, theirs = gpu.parse_own(text), gpu.parse(text)
    assert [gpu.print_genome(mine, i) for i in range(mine.n)] == [gpu.print_genome(theirs, i) for i in range(theirs.n)]
    assert gpu.print_genome(mine, 1) == gpu.print_genome(g, 1) and gpu.print_genome(mine, 2) == gpu.print_genome(g, 2)


def _interp_pairs():
    fx = [f for f in _fixtures() if os.sep + "c5" + os.sep in f]
    # neighbours in the sorted list: unrelated random genomes with different numbers of xforms,
    # final xforms on one side only, all 99 variations between them
    return [(fx[i], fx[(i * 7 + 3) % len(fx)], fx[(i * 11 + 5) % len(fx)], fx[(i * 13 + 1) % len(fx)]) for i in range(0, len(fx), 3)]


@pytest.mark.parametrize("quad", _interp_pairs(), ids=lambda q: os.path.basename(q[0])[:-6])
def test_own_interpolation_equals_flam3_interpolate(gpu, quad):
    """f3b_interpolate (f3b_frontend.c: own flam3_align / flam3_interpolate_n / Catmull-Rom) against
    the reference's flam3_interpolate on four unrelated random genomes as key frames: padding
    xforms and their 'better identities', parameters borrowed from the neighbour in time,
    log-polar affine parts, palettes - linear segments and a smooth one, inside segments, on
    and beyond the key frames, with and without stagger. Byte for byte (genome struct, xform arrays, chaos matrix)."""
    import ctypes as C
    from types import SimpleNamespace
    import numpy as np
    from flam3_b200.genome_xml import set_flame_attrs
    flames = []
    for k, path in enumerate(quad):
        x = open(path).read()
        x = re.sub(r'\btime="[^"]*"', "", x, count=1).replace("<flame ", '<flame time="%d" ' % (k * 40), 1)
        if k == 1:
            x = set_flame_attrs(x, interpolation="smooth")
        flames.append(x)
    g = gpu.parse("<flames>" + "".join(flames) + "</flames>")
    assert g.n == 4
    d = gpu._dll
    size = gpu.abi_sizes()["flam3_genome"]
    d.flam3_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    d.flam3_interpolate.restype = None
    d.f3b_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    d.clear_cp.argtypes = [C.c_void_p, C.c_int]
    for t, stagger in ((-5.0, 0.0), (0.0, 0.0), (13.5, 0.0), (40.0, 0.0), (47.25, 0.0), (79.9, 0.0), (80.0, 0.0), (101.0, 0.0),
                       (120.0, 0.0), (130.0, 0.0), (13.5, 0.3), (29.0, 1.0), (95.0, 0.65), (47.25, 0.5), (3.0, 0.9)):
        ra, rb = (C.c_char * size)(), (C.c_char * size)()
        d.flam3_interpolate(g.ptr, 4, t, stagger, ra)
        assert d.f3b_interpolate(g.ptr, 4, t, stagger, rb) == 0
        (ga, xa, ca), = _raw(gpu, SimpleNamespace(ptr=C.addressof(ra), n=1))
        (gb, xb, cb), = _raw(gpu, SimpleNamespace(ptr=C.addressof(rb), n=1))
        diff = [i for i in range(len(ga)) if ga[i] != gb[i]]
        assert not diff, "t=%g stagger=%g: genome bytes differ at %s" % (t, stagger, diff[:12])
        assert ca == cb, t
        bad = np.where((xa != xb).any(axis=0))[0]
        assert bad.size == 0, "t=%g stagger=%g: xform bytes differ at offsets %s" % (t, stagger, bad[:16])
        d.clear_cp(ra, 0)
        d.clear_cp(rb, 0)


MOTION = """
<flame time="0" size="64 48" center="0 0" scale="20">
   <xform weight="0.5" color="0.3" coefs="1 0.1 -0.1 1 0.2 0.3" linear="0.5" julian="0.4" julian_power="3" julian_dist="0.8">
      <motion motion_frequency="2" motion_function="sin" color="0.1" julian_power="1.5" coefs="0 0 0 0 0.3 0"/>
      <motion motion_frequency="1" motion_function="hill" linear="0.25" oscilloscope_frequency="0.5" oscope_damping="0.125"
              separation_x="0.2" post="0.1 0 0 0.1 0 0" color_speed="0.05"/>
   </xform>
   <xform weight="0.5" color="1" coefs="0.5 0 0 0.5 0.5 0" spherical="1" oscilloscope="0.2"/>
   <finalxform color="0" coefs="1 0 0 1 0 0" linear="1">
      <motion motion_frequency="3" motion_function="triangle" coefs="0 0.2 -0.2 0 0 0"/>
   </finalxform>
   <color index="0" rgb="255 0 0"/><color index="255" rgb="0 0 255"/>
</flame>
"""


def test_motion_elements(gpu):
    """<motion> children of xforms (parser.c:763-777; they act in flam3-genome's sequences only):
    read into the same bytes as the reference's parser, written like the reference's printer
    (field names, %f, only what is not zero), and the text reads back alike."""
    import ctypes as C
    import numpy as np
    a, b = _raw(gpu, gpu.parse(MOTION)), _raw(gpu, gpu.parse_own(MOTION))
    assert a[0][0] == b[0][0] and a[0][2] == b[0][2]
    assert a[0][1].shape == b[0][1].shape and a[0][1].shape[0] == 3 + 3 and np.array_equal(a[0][1], b[0][1])
    d = gpu._dll
    d.flam3_version.restype = C.c_char_p
    d.f3b_print_flam3.restype = C.c_void_p
    d.f3b_print_flam3.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    g = gpu.parse(MOTION)
    want = gpu.print_genome(g, 0)
    got = C.string_at(d.f3b_print_flam3(g.ptr, b"FLAM3-" + d.flam3_version(), None)).decode()
    assert "<motion motion_frequency" in want and got == want
    again_a, again_b = _raw(gpu, gpu.parse(got)), _raw(gpu, gpu.parse_own(got))
    assert np.array_equal(again_a[0][1], again_b[0][1]) and again_a[0][0] == again_b[0][0]
    for bad in ('<flame size="10 10"><xform weight="1" coefs="1 0 0 1 0 0" motion_frequency="2"/></flame>',
                '<flame size="10 10"><xform weight="1" coefs="1 0 0 1 0 0"><motion motion_function="saw"/></xform></flame>',
                '<flame size="10 10"><xform weight="1" coefs="1 0 0 1 0 0"><motion motion_function="sin" chaos="1 2"/></xform></flame>'):
        with pytest.raises(ValueError):
            gpu.parse_own(bad)


def test_random_symmetry_follows_the_c_library(gpu):
    """<symmetry kind="0"/>: a symmetry drawn from random(). From the same state of the C
    library's generator both parsers draw the same one and add the same xforms."""
    import ctypes as C
    libc = C.CDLL(None)
    xml = '<flame size="10 10"><symmetry kind="0"/><xform weight="1" color="0.3" coefs="0.5 0 0 0.5 0.1 0" linear="1"/></flame>'
    kinds = set()
    for seed in range(1, 40):
        libc.srandom(seed)
        a = gpu.parse(xml)
        libc.srandom(seed)
        b = gpu.parse_own(xml)
        ta, tb = gpu.print_genome(a, 0), gpu.print_genome(b, 0)
        assert ta == tb, seed
        m = re.search(r'<symmetry kind="(-?\d+)"', ta)
        kinds.add(int(m.group(1)) if m else 1)
    assert len(kinds) >= 6, kinds


def test_own_parser_refuses_what_it_does_not_read(gpu):
    for bad in (
                '<flame size="10 10"><palette index0="1" count="256" format="RGB"/><xform weight="1" coefs="1 0 0 1 0 0"/></flame>',
                '<flame size="10 oops"><xform weight="1" coefs="1 0 0 1 0 0"/></flame>',
                '<flame size="10 10"><xform weight="1x" coefs="1 0 0 1 0 0"/></flame>'):
        with pytest.raises(ValueError):
            gpu.parse_own(bad)
