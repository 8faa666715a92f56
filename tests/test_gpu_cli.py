"""The reference's own command-line programs, compiled UNMODIFIED from its sources and
linked against the drop-in libflam3 (dropin/_build), run on the GPU box: README.txt:42-51's
smoke test (`flam3-render < test.flam3`, `flam3-animate`) with the B200 path underneath, next
to the same program linked against the unmodified library (oracle/_ref/flam3-render-ref).
"""
import os
import subprocess

import numpy as np
import pytest

from imgio import read_ppm
from parity_util import psnr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "test_both.flam3")


def run_cli(exe, tmp, prefix, extra_env=None, stdin=FIXTURE):
    env = dict(os.environ)
    env.update({"format": "ppm", "isaac_seed": "flam3b200", "prefix": os.path.join(tmp, prefix),
                "verbose": "0"})
    env.update(extra_env or {})
    with open(stdin) as f:
        p = subprocess.run([exe], stdin=f, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p


def test_flam3_render_cli_both_flames(tmp_path):
    tmp = str(tmp_path)
    run_cli(os.path.join(ROOT, "dropin", "_build", "flam3-render"), tmp, "gpu_")
    run_cli(os.path.join(ROOT, "oracle", "_ref", "flam3-render-ref"), tmp, "refa_",
            {"nthreads": "1"})
    run_cli(os.path.join(ROOT, "oracle", "_ref", "flam3-render-ref"), tmp, "refb_",
            {"nthreads": "1", "isaac_seed": "another"})
    for i in range(2):
        g = read_ppm(os.path.join(tmp, "gpu_%05d.ppm" % i))
        a = read_ppm(os.path.join(tmp, "refa_%05d.ppm" % i))
        b = read_ppm(os.path.join(tmp, "refb_%05d.ppm" % i))
        assert g.shape == (480, 640, 3)
        assert min(psnr(g, a), psnr(g, b)) >= psnr(a, b) - 1.5


def test_flam3_render_cli_strips(tmp_path):
    """nstrips is implemented by the CALLER (flam3-render.c:311-348) on top of flam3_render:
    height/center/sample_density are rewritten per strip and f.rc is restored. It must
    keep working with the drop-in underneath."""
    tmp = str(tmp_path)
    run_cli(os.path.join(ROOT, "dropin", "_build", "flam3-render"), tmp, "s1_")
    run_cli(os.path.join(ROOT, "dropin", "_build", "flam3-render"), tmp, "s4_", {"nstrips": "4"})
    a = read_ppm(os.path.join(tmp, "s1_00000.ppm"))
    b = read_ppm(os.path.join(tmp, "s4_00000.ppm"))
    assert a.shape == b.shape
    assert psnr(a, b) > 30.0


def test_flam3_render_cli_png_rgba16(tmp_path):
    tmp = str(tmp_path)
    run_cli(os.path.join(ROOT, "dropin", "_build", "flam3-render"), tmp, "p_",
            {"format": "png", "transparency": "1", "bpc": "16", "qs": "0.5"})
    data = open(os.path.join(tmp, "p_00000.png"), "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    # IHDR: 640x480, 16 bit, colour type 6
    assert data[16:26] == (640).to_bytes(4, "big") + (480).to_bytes(4, "big") + bytes([16, 6])


def test_flam3_animate_cli_frames(tmp_path):
    """flam3-animate interpolates between the two key frames with the default 1000
    temporal samples per frame (motion blur), all in one chaos-game launch per frame."""
    tmp = str(tmp_path)
    run_cli(os.path.join(ROOT, "dropin", "_build", "flam3-animate"), tmp, "an_",
            {"begin": "40", "end": "42"})
    frames = [read_ppm(os.path.join(tmp, "an_%05d.ppm" % t)) for t in (40, 41, 42)]
    for f in frames:
        assert f.shape == (480, 640, 3) and f.mean() > 5
    # consecutive frames of a smooth interpolation are close but not identical
    assert 20.0 < psnr(frames[0], frames[1]) < 60.0


def test_flam3_animate_cli_matches_the_reference_cli(tmp_path):
    """Config C4's two sides, frame by frame: the unmodified flam3-animate linked against the
    drop-in against the same program linked against the unmodified library (1000 temporal
    samples per frame, motion blur). Bound as everywhere: as close to a reference render as a
    second reference render with another seed is, less 1.5 dB."""
    tmp = str(tmp_path)
    times = {"begin": "20", "end": "21"}
    run_cli(os.path.join(ROOT, "dropin", "_build", "flam3-animate"), tmp, "g_", times)
    run_cli(os.path.join(ROOT, "oracle", "_ref", "flam3-animate-ref"), tmp, "ra_", dict(times, nthreads="8"))
    run_cli(os.path.join(ROOT, "oracle", "_ref", "flam3-animate-ref"), tmp, "rb_",
            dict(times, nthreads="8", isaac_seed="another"))
    for t in (20, 21):
        g, a, b = (read_ppm(os.path.join(tmp, "%s%05d.ppm" % (p, t))) for p in ("g_", "ra_", "rb_"))
        assert g.shape == a.shape == (480, 640, 3)
        assert min(psnr(g, a), psnr(g, b)) >= psnr(a, b) - 1.5, (t, psnr(g, a), psnr(g, b), psnr(a, b))
        assert abs(g.mean() - a.mean()) <= max(0.5, 0.02 * a.mean())


def test_animate_frame_scheduler_round_robin(tmp_path):
    """flam3_b200.animate deals the frames of one flam3-animate job
    round-robin to worker processes (one per GPU; here two workers on the one GPU). Every
    frame appears once under the reference's file name and matches the single-process
    frame up to Monte-Carlo noise (the seeds differ, see flam3_b200/animate.py)."""
    import sys
    tmp = str(tmp_path)
    run_cli(os.path.join(ROOT, "dropin", "_build", "flam3-animate"), tmp, "one_",
            {"begin": "30", "end": "34"})
    env = dict(os.environ)
    env.update({"format": "ppm", "isaac_seed": "flam3b200", "prefix": os.path.join(tmp, "rr_"),
                "verbose": "0", "in": FIXTURE, "begin": "30", "end": "34", "PYTHONPATH": ROOT})
    p = subprocess.run([sys.executable, "-m", "flam3_b200.animate", "--gpus", "2", "--devices", "0,0"],
                       env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    for t in range(30, 35):
        a = read_ppm(os.path.join(tmp, "one_%05d.ppm" % t))
        b = read_ppm(os.path.join(tmp, "rr_%05d.ppm" % t))
        assert a.shape == b.shape == (480, 640, 3)
        assert psnr(a, b) > 28.0, (t, psnr(a, b))
