"""SURVEY §8(f) row 4, end to end: dropin/_build/f3b-render is a flame renderer made of this
repository's code only - own genome/palette reader and image writers (f3b_frontend.c), the
flam3_render drop-in, libflam3_b200.so - with NO reference object linked. For the same genome
file and isaac_seed it must write the very bytes the reference's unmodified flam3-render writes
when linked against the drop-in (same genomes in, same frame seeding, same render)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OWN = os.path.join(ROOT, "dropin", "_build", "f3b-render")
REF_CLI = os.path.join(ROOT, "dropin", "_build", "flam3-render")


def _run(exe, tmp, prefix, fixture, **env_extra):
    env = dict(os.environ)
    env.update({"format": "ppm", "isaac_seed": "frontend", "prefix": os.path.join(tmp, prefix), "verbose": "0"})
    env.update(env_extra)
    with open(fixture) as f:
        p = subprocess.run([exe], stdin=f, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]


def test_standalone_renderer_links_no_reference_code():
    syms = subprocess.run(["nm", "-D", "--defined-only", OWN], capture_output=True, text=True).stdout + \
        subprocess.run(["nm", "--defined-only", OWN], capture_output=True, text=True).stdout
    defined = {line.split()[-1] for line in syms.splitlines() if len(line.split()) >= 3}
    assert "flam3_render" in defined and "f3b_parse_flam3" in defined
    # (f3b_render.c refers to flam3_align / flam3_interpolate_n / rgb2hsv WEAKLY, for the cached
    # interpolation between key frames of the drop-in; here they stay undefined and unused)
    for name in ("flam3_parse_xml2", "parse_flame_element", "flam3_get_palette", "flam3_interpolate_n",
                 "flam3_align", "flam3_print", "iter_thread", "de_thread", "rgb2hsv", "flam3_add_xforms"):
        assert name not in defined, name
    libs = subprocess.run(["ldd", OWN], capture_output=True, text=True).stdout
    assert "libflam3_b200.so" in libs and "libflam3.so" not in libs and "flam3_ref" not in libs


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,extra", [("test_both.flam3", {}),
                                           ("c5/c5_mix101.flam3", {"qs": "0.05"}),
                                           ("c5/c5_var33.flam3", {"qs": "0.05", "ss": "0.5", "nthreads": "4"}),
                                           ("test_flame1.flam3", {"qs": "2", "nstrips": "7"})])
def test_standalone_renderer_writes_the_same_bytes(tmp_path, fixture, extra):
    tmp = str(tmp_path)
    path = os.path.join(ROOT, "tests", "golden", fixture)
    _run(OWN, tmp, "own_", path, **extra)
    _run(REF_CLI, tmp, "cli_", path, **extra)
    own = sorted(f for f in os.listdir(tmp) if f.startswith("own_"))
    cli = sorted(f for f in os.listdir(tmp) if f.startswith("cli_"))
    assert own and len(own) == len(cli)
    for a, b in zip(own, cli):
        da, db = open(os.path.join(tmp, a), "rb").read(), open(os.path.join(tmp, b), "rb").read()
        assert len(da) > 1000 and da == db, (a, b)


@pytest.mark.gpu
def test_standalone_animation_writes_the_same_bytes(tmp_path, flame0, flame1):
    """Motion-blurred frames between two key frames: f3b-render in its flam3-animate mode (own
    interpolation: f3b_interpolate for every temporal sample) against the unmodified
    flam3-animate linked to the drop-in (the reference's flam3_align / flam3_interpolate_n through
    the cached path). The interpolated genomes are equal byte for byte
    (tests/test_cpu_frontend.py), so the images are."""
    from flam3_b200.genome_xml import resize
    tmp = str(tmp_path)
    a = resize(flame0, 160, 120, scale=60, quality=20, temporal_samples=24)
    b = resize(flame1, 160, 120, scale=60, quality=20, temporal_samples=24).replace('time="0"', 'time="10"', 1)
    path = os.path.join(tmp, "anim.flam3")
    with open(path, "w") as f:
        f.write("<flames>" + a + b + "</flames>")
    _run(OWN, tmp, "own_", path, begin="3", end="4", write_genome="1")
    _run(os.path.join(ROOT, "dropin", "_build", "flam3-animate"), tmp, "cli_", path, begin="3", end="4")
    own = sorted(f for f in os.listdir(tmp) if f.startswith("own_") and f.endswith(".ppm"))
    cli = sorted(f for f in os.listdir(tmp) if f.startswith("cli_"))
    assert own == ["own_00003.ppm", "own_00004.ppm"] and len(cli) == 2
    # write_genome: the frame's interpolated genome beside the image (flam3-animate.c:264-279)
    text = open(os.path.join(tmp, "own_00003.ppm.flam3")).read()
    assert text.startswith("<flame version=") and 'time="3"' in text and text.rstrip().endswith("</flame>")
    for x, y in zip(own, cli):
        da, db = open(os.path.join(tmp, x), "rb").read(), open(os.path.join(tmp, y), "rb").read()
        assert len(da) > 1000 and da == db, (x, y)


@pytest.mark.gpu
def test_standalone_renderer_png(tmp_path):
    tmp = str(tmp_path)
    _run(OWN, tmp, "p_", os.path.join(ROOT, "tests", "golden", "test_flame0.flam3"),
         format="png", bpc="16", transparency="1", qs="0.5")
    data = open(os.path.join(tmp, "p_00000.png"), "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    assert data[16:26] == (640).to_bytes(4, "big") + (480).to_bytes(4, "big") + bytes([16, 6])
    # the genome travels in a text chunk, as with flam3-render (png.c:64-90)
    assert b"<flame version=" in data and b"</flame>" in data and b"temporal_samples=" in data
