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


def test_own_parser_refuses_what_it_does_not_read(gpu):
    for bad in ('<flame size="10 10"><symmetry kind="0"/><xform weight="1" coefs="1 0 0 1 0 0" linear="1"/></flame>',
                '<flame size="10 10"><xform weight="1" coefs="1 0 0 1 0 0" linear="1"><motion linear="1"/></xform></flame>',
                '<flame size="10 10"><palette index0="1" index1="2" blend="0.5"/><xform weight="1" coefs="1 0 0 1 0 0"/></flame>',
                '<flame size="10 oops"><xform weight="1" coefs="1 0 0 1 0 0"/></flame>',
                '<flame size="10 10"><xform weight="1x" coefs="1 0 0 1 0 0"/></flame>'):
        with pytest.raises(ValueError):
            gpu.parse_own(bad)
