"""Frame scheduler for flam3-animate across the GPUs of one node.

flam3-animate renders frames `begin, begin+dtime, ... <= end` one after the other
(flam3-animate.c:225-327); frames are independent, so they are dealt round-robin to one
worker process per GPU. Each worker is the UNMODIFIED flam3-animate binary linked against
the drop-in libflam3.so, told which frames are its own through the reference's own
environment interface: `begin = first + rank*dtime`, `dtime = n*dtime`, same `end`. Every
frame is its own file (`<prefix>%05d.<format>`, flam3-animate.c:257-258), so no ordering
between writers is needed.

    env in=anim.flam3 prefix=out/ begin=0 end=99 python -m flam3_b200.animate --gpus 8

Note: like any multi-process split of the reference (Electric Sheep renders one frame per
process), a frame's random seed differs from the one it would get in a single process that
rendered all earlier frames first; the images are statistically equivalent, not identical.
"""
import argparse
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_BINARY = os.path.join(ROOT, "dropin", "_build", "flam3-animate")


def plan_frames(first, last, dtime, nworkers):
    """[(begin, dtime)] per worker, or None for a worker without frames: worker r renders
    first + r*dtime, first + (r + nworkers)*dtime, ... <= last."""
    if dtime < 1:
        raise ValueError("dtime must be positive, not %d." % dtime)
    plan = []
    for r in range(nworkers):
        b = first + r * dtime
        plan.append((b, nworkers * dtime) if b <= last else None)
    return plan


def frames_of(plan, last):
    """The frame times each worker of `plan` renders (for tests)."""
    return [[] if p is None else list(range(p[0], last + 1, p[1])) for p in plan]


def frame_range(env, xml_text):
    """first/last frame as flam3-animate derives them (flam3-animate.c:99-116, 179-187)."""
    if "frame" in env or "time" in env:
        t = int(env.get("frame", env.get("time")))
        return t, t
    times = [float(t) for t in re.findall(r"<flame\b[^>]*?\btime=\"([^\"]+)\"", xml_text)]
    first = int(env["begin"]) if "begin" in env else (int(times[0]) if times else 0)
    if "end" in env:
        last = int(env["end"])
    else:
        last = (int(times[-1]) - 1) if times else 0
        if last < first:
            last = first
    return first, last


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--gpus", type=int, default=1, help="worker processes, one per GPU")
    ap.add_argument("--devices", default=None, help="comma list of CUDA device indices (default 0..gpus-1)")
    ap.add_argument("--binary", default=DEFAULT_BINARY)
    args = ap.parse_args(argv)
    env = dict(os.environ)
    if not os.path.exists(args.binary):
        sys.stderr.write("animate: %s is missing (run `make dropin`)\n" % args.binary)
        return 1
    src = env.get("in")
    xml = open(src).read() if src else sys.stdin.read()
    first, last = frame_range(env, xml)
    dtime = int(env.get("dtime", "1"))
    devices = [int(d) for d in args.devices.split(",")] if args.devices else list(range(args.gpus))
    plan = plan_frames(first, last, dtime, len(devices))
    procs = []
    for dev, p in zip(devices, plan):
        if p is None:
            continue
        e = dict(env)
        for k in ("frame", "time"):
            e.pop(k, None)
        e.update(begin=str(p[0]), end=str(last), dtime=str(p[1]), F3B_DEVICE=str(dev))
        procs.append(subprocess.Popen([args.binary], env=e,
                                      stdin=subprocess.PIPE if not src else None))
        if not src:
            procs[-1].stdin.write(xml.encode())
            procs[-1].stdin.close()
    rc = 0
    for p in procs:
        rc = max(rc, p.wait())
    return rc


if __name__ == "__main__":
    sys.exit(main())
