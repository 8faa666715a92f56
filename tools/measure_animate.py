"""Config C4 end to end through the reference's own CLI: `flam3-animate` over the two key frames
of test.flam3, 640x480 quality 10, 1000 temporal samples per frame, PPM files written.
GPU: dropin/_build/flam3-animate, 100 frames, per-frame times from F3B_TIMING (time inside
flam3_render + time in the unmodified CLI between calls); CPU: oracle/_ref/flam3-animate-ref,
5 frames, wall clock."""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "test_both.flam3")
def run(exe, begin, end, extra=None):
    tmp = tempfile.mkdtemp()
    env = dict(os.environ)
    env.update({"format": "ppm", "isaac_seed": "c4", "prefix": tmp + "/f_", "verbose": "0",
                "begin": str(begin), "end": str(end)})
    env.update(extra or {})
    t = time.perf_counter()
    with open(FIX) as f:
        p = subprocess.run([exe], stdin=f, env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t
    assert p.returncode == 0, p.stderr[-500:]
    n = len([x for x in os.listdir(tmp) if x.endswith(".ppm")])
    return dt, n, p.stderr
gpu = os.path.join(ROOT, "dropin", "_build", "flam3-animate")
ref = os.path.join(ROOT, "oracle", "_ref", "flam3-animate-ref")
dt, n, err = run(gpu, 0, 99, {"F3B_TIMING": "1"})
fr = [(float(a), float(b)) for a, b in re.findall(r"flam3_render ([0-9.]+) ms, ([0-9.]+) ms in the caller", err)]
steady = fr[1:]
per = sum(a + b for a, b in steady) / len(steady)
print("GPU  flam3-animate: %d frames written, process wall %.2f s; first frame (CUDA start-up) %.0f ms; "
      "steady state %.2f ms/frame (%.2f in flam3_render + %.2f in the CLI) = %.0f frames/s"
      % (n, dt, fr[0][0], per, sum(a for a, _ in steady) / len(steady), sum(b for _, b in steady) / len(steady), 1e3 / per))
tr, nr, _ = run(ref, 0, 4)
print("CPU  flam3-animate (unmodified, %d cores): %d frames %.2f s -> %.0f ms/frame" % (os.cpu_count(), nr, tr, 1e3 * tr / nr))
