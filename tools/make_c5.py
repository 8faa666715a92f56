"""Generates the random genomes of config C5 with the UNMODIFIED reference
flam3-genome (oracle/_ref/flam3-genome-ref, built by `make -C oracle`):

    for v in 0..98:  env seed=$v isaac_seed=v$v use_vars=$v template=<1024x1024 q1000> flam3-genome
    for s in 100..123: the same without use_vars (random mixes of variations, symmetries, ...)

Run in the build container (needs /root/reference for the palette file); the outputs are
committed under tests/golden/c5/ because /root/reference does not exist on the GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "flam3-genome-ref")
REF = os.environ.get("F3_REF", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "c5")
TEMPLATE = ('<flame time="0" size="1024 1024" center="0 0" scale="256" quality="1000" supersample="1" '
            'temporal_samples="1" estimator_radius="9"/>\n')
os.makedirs(OUT, exist_ok=True)
tmpl = os.path.join(OUT, "_template.xml")
open(tmpl, "w").write(TEMPLATE)
jobs = [("var%02d" % v, {"seed": str(v), "isaac_seed": "v%d" % v, "use_vars": str(v)}) for v in range(99)]
jobs += [("mix%03d" % s, {"seed": str(s), "isaac_seed": "m%d" % s}) for s in range(100, 124)]
for name, env in jobs:
    e = dict(os.environ)
    e.update(env)
    e.update({"template": tmpl, "flam3_palettes": os.path.join(REF, "flam3-palettes.xml"), "verbose": "0"})
    p = subprocess.run([EXE], env=e, capture_output=True, text=True, timeout=600)
    if p.returncode != 0 or "<flame" not in p.stdout:
        # e.g. use_vars=67 (pre_blur alone never makes a picture: flam3-genome gives up)
        print("skipped %s: rc=%d %s" % (name, p.returncode, p.stderr[-200:].strip()), file=sys.stderr)
        continue
    # keep the <flame> element only (flam3-genome wraps it in <pick>)
    a, b = p.stdout.index("<flame"), p.stdout.rindex("</flame>") + len("</flame>")
    open(os.path.join(OUT, "c5_%s.flam3" % name), "w").write(p.stdout[a:b] + "\n")
    print(name, len(p.stdout), file=sys.stderr)
