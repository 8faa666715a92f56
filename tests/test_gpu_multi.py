"""Several GPUs on one frame, inside the C library (north_star: "the host side
stays C"). F3B_DEVICES spreads every flam3_render call over the listed devices: each runs its
share of the chains into its own buckets, then owns a strip of image rows, sums all ranks'
buckets of the rows that strip needs straight out of the peers' memory (saturating, halo
rows included) and writes its image rows into the first device's image. An entry may repeat,
so the whole path runs here with two or three contexts on ONE GPU; because the buckets are
exact integers, the image must equal the single-context image byte for byte."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

SCRIPT = r"""
import os, sys, numpy as np
sys.path.insert(0, %(root)r)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize, set_flame_attrs
gpu = fb.load_dropin()
f0 = open(os.path.join(%(root)r, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(%(root)r, "tests/golden/test_flame1.flam3")).read()
cases = {
  "os1": (resize(f0, 640, 363, scale=160, quality=50, temporal_samples=2), {}),
  "os2_earlyclip": (resize(f0, 320, 200, scale=80, quality=50, oversample=2, temporal_samples=1), {"earlyclip": 1}),
  "os3_rgba16": (resize(f0, 200, 150, scale=50, quality=30, oversample=3, temporal_samples=1),
                 {"nchan": 4, "transp": 1, "bytes_per_channel": 2}),
  "motion_blur_batches": ("<flames>" + set_flame_attrs(resize(f0, 256, 192, scale=64, quality=40, temporal_samples=20), passes=2)
                          + set_flame_attrs(resize(f1, 256, 192, scale=64, quality=40, temporal_samples=20), passes=2).replace('time="0"', 'time="100"')
                          + "</flames>", {"ngenomes": 2, "time": 50.0}),
  "linear_palette_count_plane": (set_flame_attrs(resize(f0, 300, 200, scale=75, quality=40, temporal_samples=1), palette_mode="linear"), {}),
  "no_de": (resize(f0, 300, 200, scale=75, quality=40, temporal_samples=1, estimator_radius=0), {}),
  "tiny_more_ranks_than_needed": (resize(f0, 64, 5, scale=16, quality=40, temporal_samples=1), {}),
}
out = {}
for name, (xml, kw) in cases.items():
    g = gpu.parse(xml)
    img, st, rc = gpu.render(g, isaac_seed="multi", sub_batch_size=300, **kw)
    assert st["rv"] == 0, (name, st)
    out[name] = img
    out[name + "_iters"] = np.array([st["num_iters"], st["badvals"]])
    out[name + "_rc"] = np.frombuffer(rc, dtype=np.uint8)
np.savez(sys.argv[1], **out)
"""


def _run(devices, path):
    env = dict(os.environ)
    env.pop("F3B_DEVICE", None)
    if devices:
        env["F3B_DEVICES"] = devices
    else:
        env.pop("F3B_DEVICES", None)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, path], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(path)


@pytest.mark.parametrize("devices", ["0,0", "0,0,0", "0,0,0,0,0"])
def test_frame_spread_over_contexts_is_byte_identical(tmp_path, devices):
    one = _run(None, str(tmp_path / "one.npz"))
    many = _run(devices, str(tmp_path / "many.npz"))
    assert set(one.files) == set(many.files)
    for k in one.files:
        assert one[k].shape == many[k].shape, k
        if k.endswith("_iters"):
            assert one[k][0] == many[k][0], k
            continue
        if not k.endswith("_rc"):
            assert one[k].max() > 0, k
        assert np.array_equal(one[k], many[k]), "%s: %d of %d bytes differ" % (
            k, int((one[k] != many[k]).sum()), one[k].size)


def test_unmodified_flam3_render_uses_f3b_devices(tmp_path):
    """The reference's own CLI, linked against the drop-in, spreads its frames over the devices
    of F3B_DEVICES without knowing (flam3-render.c:345 just calls flam3_render)."""
    exe = os.path.join(ROOT, "dropin", "_build", "flam3-render")
    fixture = os.path.join(ROOT, "tests", "golden", "test_both.flam3")
    imgs = []
    for tag, devices in (("one_", None), ("three_", "0,0,0")):
        env = dict(os.environ)
        env.update({"format": "ppm", "isaac_seed": "cli-multi", "prefix": str(tmp_path / tag),
                    "verbose": "0", "nthreads": "4"})
        env.pop("F3B_DEVICES", None)
        if devices:
            env["F3B_DEVICES"] = devices
        with open(fixture) as f:
            r = subprocess.run([exe], stdin=f, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        imgs.append([open(tmp_path / ("%s%05d.ppm" % (tag, i)), "rb").read() for i in range(2)])
    for a, b in zip(*imgs):
        assert len(a) > 640 * 480 * 3 and a == b


THREADS_SCRIPT = r"""
import os, sys, threading, numpy as np
sys.path.insert(0, %(root)r)
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
gpu = fb.load_dropin()
f0 = open(os.path.join(%(root)r, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(%(root)r, "tests/golden/test_flame1.flam3")).read()
jobs = [(resize(f0, 400, 300, scale=100, quality=300, temporal_samples=1), "t0"),
        (resize(f1, 320, 240, scale=80, quality=400, temporal_samples=1), "t1"),
        (resize(f0, 256, 256, scale=64, quality=200, oversample=2, temporal_samples=1), "t2"),
        (resize(f1, 500, 200, scale=120, quality=300, temporal_samples=3), "t3")]
genomes = [gpu.parse(x) for x, _ in jobs]
seq = [gpu.render(g, isaac_seed=s)[0] for g, (_, s) in zip(genomes, jobs)]
par = [None] * len(jobs)
def work(i):
    for _ in range(3):
        par[i] = gpu.render(genomes[i], isaac_seed=jobs[i][1])[0]
th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
[t.start() for t in th]
[t.join() for t in th]
ok = all(np.array_equal(a, b) and a.max() > 0 for a, b in zip(seq, par))
print("concurrent == sequential:", ok)
sys.exit(0 if ok else 1)
"""


@pytest.mark.parametrize("contexts", ["1", "3"])
def test_concurrent_frames_from_several_host_threads(contexts):
    """SURVEY §8(b) Threading: independent frames may be rendered concurrently from different
    host threads. Four threads render four different frames three times each; with
    F3B_CONTEXTS=3 three of them are on the GPU at once (each call takes a free slot: context,
    batch buffer, filter cache), with 1 they take turns. Either way every image equals the one
    rendered alone."""
    env = dict(os.environ, F3B_CONTEXTS=contexts)
    env.pop("F3B_DEVICES", None)
    r = subprocess.run([sys.executable, "-c", THREADS_SCRIPT % {"root": ROOT}], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-1500:])
