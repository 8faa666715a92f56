"""One single-xform genome per variation (config C5 of SURVEY.md §8(d), correctness side):
the variation under test at weight 0.8 plus a little linear, a non-trivial affine and post
transform, and every parameter of the variation set to a non-default value.

Shared by the CPU tests (C oracle vs reference) and the GPU tests (CUDA vs reference)."""
import ctypes as C

import numpy as np

VAR_NAMES = [
    "linear", "sinusoidal", "spherical", "swirl", "horseshoe", "polar", "handkerchief", "heart",
    "disc", "spiral", "hyperbolic", "diamond", "ex", "julia", "bent", "waves", "fisheye", "popcorn",
    "exponential", "power", "cosine", "rings", "fan", "blob", "pdj", "fan2", "rings2", "eyefish",
    "bubble", "cylinder", "perspective", "noise", "julian", "juliascope", "blur", "gaussian_blur",
    "radial_blur", "pie", "ngon", "curl", "rectangles", "arch", "tangent", "square", "rays", "blade",
    "secant2", "twintrian", "cross", "disc2", "super_shape", "flower", "conic", "parabola", "bent2",
    "bipolar", "boarders", "butterfly", "cell", "cpow", "curve", "edisc", "elliptic", "escher", "foci",
    "lazysusan", "loonie", "pre_blur", "modulus", "oscilloscope", "polar2", "popcorn2", "scry",
    "separation", "split", "splits", "stripes", "wedge", "wedge_julia", "wedge_sph", "whorl", "waves2",
    "exp", "log", "sin", "cos", "tan", "sec", "csc", "cot", "sinh", "cosh", "tanh", "sech", "csch",
    "coth", "auger", "flux", "mobius"]

# variations that draw random numbers (column "rng")
RANDOM_VARS = {13, 31, 32, 33, 34, 35, 36, 37, 41, 43, 44, 45, 47, 50, 51, 52, 53, 56, 59, 67, 78}

PARAMS = {
    "blob": dict(blob_low=0.3, blob_high=1.2, blob_waves=4),
    "pdj": dict(pdj_a=1.1, pdj_b=-0.7, pdj_c=1.7, pdj_d=0.6),
    "fan2": dict(fan2_x=0.6, fan2_y=0.4),
    "rings2": dict(rings2_val=0.7),
    "perspective": dict(perspective_angle=0.4, perspective_dist=2.5),
    "julian": dict(julian_power=3, julian_dist=0.9),
    "juliascope": dict(juliascope_power=-2, juliascope_dist=1.1),
    "radial_blur": dict(radial_blur_angle=0.35),
    "pie": dict(pie_slices=6, pie_rotation=0.3, pie_thickness=0.6),
    "ngon": dict(ngon_sides=5, ngon_power=3, ngon_circle=1.2, ngon_corners=2),
    "curl": dict(curl_c1=0.7, curl_c2=0.3),
    "rectangles": dict(rectangles_x=0.6, rectangles_y=0.9),
    "disc2": dict(disc2_rot=1.3, disc2_twist=7.5),
    "super_shape": dict(super_shape_rnd=0.3, super_shape_m=5, super_shape_n1=2.5,
                        super_shape_n2=1.5, super_shape_n3=0.8, super_shape_holes=0.2),
    "flower": dict(flower_petals=5, flower_holes=0.3),
    "conic": dict(conic_eccentricity=0.6, conic_holes=0.2),
    "parabola": dict(parabola_height=0.8, parabola_width=1.3),
    "bent2": dict(bent2_x=1.7, bent2_y=0.4),
    "bipolar": dict(bipolar_shift=0.4),
    "cell": dict(cell_size=0.6),
    "cpow": dict(cpow_r=1.2, cpow_i=0.3, cpow_power=2),
    "curve": dict(curve_xamp=0.4, curve_yamp=0.7, curve_xlength=0.9, curve_ylength=1.2),
    "escher": dict(escher_beta=0.6),
    "lazysusan": dict(lazysusan_spin=0.8, lazysusan_space=0.3, lazysusan_twist=0.5,
                      lazysusan_x=0.2, lazysusan_y=-0.1),
    "modulus": dict(modulus_x=0.7, modulus_y=0.9),
    "oscilloscope": dict(oscope_separation=0.8, oscope_frequency=1.3, oscope_amplitude=0.7,
                         oscope_damping=0.4),
    "popcorn2": dict(popcorn2_x=0.3, popcorn2_y=0.5, popcorn2_c=2.2),
    "separation": dict(separation_x=0.6, separation_xinside=0.2, separation_y=0.8,
                       separation_yinside=-0.3),
    "split": dict(split_xsize=0.7, split_ysize=1.1),
    "splits": dict(splits_x=0.3, splits_y=0.5),
    "stripes": dict(stripes_space=0.4, stripes_warp=0.6),
    "wedge": dict(wedge_angle=0.5, wedge_hole=0.2, wedge_count=3, wedge_swirl=0.3),
    "wedge_julia": dict(wedge_julia_angle=0.4, wedge_julia_count=2, wedge_julia_power=4,
                        wedge_julia_dist=0.8),
    "wedge_sph": dict(wedge_sph_angle=0.5, wedge_sph_count=3, wedge_sph_hole=0.1,
                      wedge_sph_swirl=0.4),
    "whorl": dict(whorl_inside=0.6, whorl_outside=0.3),
    "waves2": dict(waves2_freqx=2.1, waves2_scalex=0.4, waves2_freqy=1.7, waves2_scaley=0.6),
    "auger": dict(auger_sym=0.3, auger_weight=0.6, auger_freq=2.4, auger_scale=0.5),
    "flux": dict(flux_spread=0.4),
    "mobius": dict(mobius_re_a=0.8, mobius_im_a=0.1, mobius_re_b=0.2, mobius_im_b=-0.3,
                   mobius_re_c=0.1, mobius_im_c=0.4, mobius_re_d=1.1, mobius_im_d=0.2),
}

COEFS = (0.72, -0.31, 0.28, 0.65, 0.12, -0.18)   # c00 c01 c10 c11 c20 c21 (the XML order)
POST = (0.95, 0.05, -0.04, 1.02, 0.03, -0.02)
WEIGHT = 0.8
LINEAR = 0.15


def xform_attrs(vid):
    name = VAR_NAMES[vid]
    a = {"weight": 1, "color": 0.3, "color_speed": 0.4, name: WEIGHT}
    if vid != 0:
        a["linear"] = LINEAR
    a.update(PARAMS.get(name, {}))
    return a


def genome_xml(vid, size=96, quality=40):
    a = xform_attrs(vid)
    attrs = " ".join('%s="%r"' % (k, float(v)) for k, v in a.items())
    pal = "\n".join('<color index="%d" rgb="%d %d %d"/>' % (i, 255 - i // 2, 64 + i // 2, (i * 5) % 256)
                    for i in range(256))
    return ('<flame time="0" size="%d %d" center="0 0" scale="%g" quality="%g" temporal_samples="1" '
            'brightness="4" gamma="3" vibrancy="1" supersample="1" filter="0.5" '
            'estimator_radius="4">\n'
            '<xform %s coefs="%s" post="%s"/>\n'
            # a plain rotating contraction as second map, so the pair has a proper attractor
            # instead of the fixed point a lone contractive xform collapses to
            '<xform weight="0.6" color="0.85" color_speed="0.5" linear="1" '
            'coefs="0.55 0.35 -0.35 0.55 0.4 -0.2"/>\n%s\n</flame>\n'
            % (size, size, size / 4.0, quality, attrs, " ".join(map(repr, COEFS)),
               " ".join(map(repr, POST)), pal))


def test_points(n=512, seed=1234):
    """(x, y, color) inputs: a grid-free spread over [-2, 2]^2 incl. points near the axes."""
    rs = np.random.RandomState(seed)
    pts = np.empty((n, 3))
    pts[:, 0] = rs.uniform(-2, 2, n)
    pts[:, 1] = rs.uniform(-2, 2, n)
    pts[:, 2] = rs.uniform(0, 1, n)
    pts[:8, :2] = [[0.5, 0.0], [0.0, 0.5], [-0.5, 0.0], [0.0, -0.7], [1.0, 1.0], [-1.0, 1.0],
                   [0.25, -0.25], [1.5, 0.1]]
    return pts


# ---- the C oracle's genome model (oracle/f3_oracle.c) ------------------------------------
def oracle_genome(olib, vid):
    olib.o_genome_new.restype = C.c_void_p
    olib.o_genome_new.argtypes = [C.c_int, C.c_int, C.c_int]
    olib.o_xform_set.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_double]
    g = olib.o_genome_new(1, -1, 0)
    for k, v in xform_attrs(vid).items():
        assert olib.o_xform_set(g, 0, k.encode(), float(v)) == 0, k
    names = ["c00", "c01", "c10", "c11", "c20", "c21"]
    for nme, v in zip(names, COEFS):
        assert olib.o_xform_set(g, 0, nme.encode(), v) == 0
    for nme, v in zip(names, POST):
        assert olib.o_xform_set(g, 0, ("p" + nme[1:]).encode(), v) == 0
    return g
