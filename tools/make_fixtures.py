"""Generate the committed genome fixtures under tests/golden/ from the reference.

Run HERE (needs /root/reference and oracle/_ref/libflam3_ref.so). The reference's own
parser + palette library turn test.flam3 (palettes referenced by number) into
self-contained genomes with explicit palettes, printed by the reference's flam3_print, so
the GPU box — which has no /root/reference and no flam3-palettes.xml — can parse them.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flam3_b200 as fb  # noqa: E402

REF = "/root/reference"


def main():
    os.environ["flam3_palettes"] = os.path.join(REF, "flam3-palettes.xml")
    ref = fb.Flam3Lib(os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so"))
    g = ref.parse(open(os.path.join(REF, "test.flam3")).read())
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    flames = [ref.print_genome(g, i) for i in range(g.n)]
    for i, s in enumerate(flames):
        with open(os.path.join(out, "test_flame%d.flam3" % i), "w") as f:
            f.write(s)
    with open(os.path.join(out, "test_both.flam3"), "w") as f:
        f.write("<test>\n" + "".join(flames) + "</test>\n")
    print("wrote", len(flames), "flames to", out)


if __name__ == "__main__":
    main()
