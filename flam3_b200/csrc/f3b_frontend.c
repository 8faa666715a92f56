/*
 * f3b_frontend.c — a self-contained front end for the render path:
 * reads .flam3 genome files and the palette library and writes images WITHOUT the reference's
 * parser.c / palettes.c / png.c, so that a program can render flames with nothing but this
 * repository's code (tools/f3b-render.c does).
 *
 *   f3b_parse_flam3        flam3_parse_xml2 + scan_for_flame_nodes + parse_flame_element +
 *                          parse_xform_xml, parser.c:220-1257
 *   f3b_get_palette        flam3_get_palette, palettes.c:25-195 (flam3-palettes.xml)
 *   f3b_genome_*           clear_cp flam3.c:1270, flam3_add_xforms flam3.c:1033 with
 *                          initialize_xforms variations.c:2401, flam3_copy flam3.c:1171
 *   f3b_write_ppm          the "ppm" branch of flam3-render.c:384-390
 *
 * It fills the reference's own `flam3_genome` (flam3.h is the public header every libflam3
 * client includes), so the genomes go straight into flam3_render. Written table-driven: one
 * table of <flame> attributes, one of <xform> parameters (attribute name = field of
 * `flam3_xform`), one of variation names. The XML reader underneath is shims/libxml/xmlshim.c
 * (this repository's mini-DOM behind the libxml2 entry points the reference uses; a real
 * libxml2 works the same).
 *
 * <edit> history is skipped silently (it does not affect rendering). Parity: for every fixture genome the reference's
 * own printer (flam3_print_to_string) gives the same text for the genome parsed here as for
 * the one parsed by the reference (tests/test_cpu_frontend.py).
 */
#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "flam3.h"

#include "../../include/flam3_b200_frontend.h"

/* ---- tables --------------------------------------------------------------------------- */
static const char *const k_variation_names[flam3_nvariations] = {
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
   "coth", "auger", "flux", "mobius"};

/* variation parameters: XML attribute = field name; value of a fresh xform (variations.c:2401) */
#define XP(field, dflt) {#field, offsetof(flam3_xform, field), dflt}
static const struct {
   const char *name;
   size_t off;
   double dflt;
} k_xform_params[] = {
   XP(blob_low, 0), XP(blob_high, 1), XP(blob_waves, 1), XP(pdj_a, 0), XP(pdj_b, 0), XP(pdj_c, 0), XP(pdj_d, 0),
   XP(fan2_x, 0), XP(fan2_y, 0), XP(rings2_val, 0), XP(perspective_angle, 0), XP(perspective_dist, 0),
   XP(julian_power, 1), XP(julian_dist, 1), XP(juliascope_power, 1), XP(juliascope_dist, 1),
   XP(radial_blur_angle, 0), XP(pie_slices, 6), XP(pie_rotation, 0), XP(pie_thickness, 0.5),
   XP(ngon_sides, 5), XP(ngon_power, 3), XP(ngon_circle, 1), XP(ngon_corners, 2), XP(curl_c1, 1), XP(curl_c2, 0),
   XP(rectangles_x, 1), XP(rectangles_y, 1), XP(amw_amp, 1), XP(disc2_rot, 0), XP(disc2_twist, 0),
   XP(super_shape_rnd, 0), XP(super_shape_m, 0), XP(super_shape_n1, 1), XP(super_shape_n2, 1),
   XP(super_shape_n3, 1), XP(super_shape_holes, 0), XP(flower_petals, 0), XP(flower_holes, 0),
   XP(conic_eccentricity, 1), XP(conic_holes, 0), XP(parabola_height, 0), XP(parabola_width, 0),
   XP(bent2_x, 1), XP(bent2_y, 1), XP(bipolar_shift, 0), XP(cell_size, 1), XP(cpow_r, 1), XP(cpow_i, 0),
   XP(cpow_power, 1), XP(curve_xamp, 0), XP(curve_yamp, 0), XP(curve_xlength, 1), XP(curve_ylength, 1),
   XP(escher_beta, 0), XP(lazysusan_space, 0), XP(lazysusan_twist, 0), XP(lazysusan_spin, 0),
   XP(lazysusan_x, 0), XP(lazysusan_y, 0), XP(modulus_x, 0), XP(modulus_y, 0), XP(oscope_separation, 1),
   XP(oscope_frequency, M_PI), XP(oscope_amplitude, 1), XP(oscope_damping, 0), XP(popcorn2_x, 0),
   XP(popcorn2_y, 0), XP(popcorn2_c, 0), XP(separation_x, 0), XP(separation_xinside, 0), XP(separation_y, 0),
   XP(separation_yinside, 0), XP(split_xsize, 0), XP(split_ysize, 0), XP(splits_x, 0), XP(splits_y, 0),
   XP(stripes_space, 0), XP(stripes_warp, 0), XP(wedge_angle, 0), XP(wedge_hole, 0), XP(wedge_count, 1),
   XP(wedge_swirl, 0), XP(wedge_julia_angle, 0), XP(wedge_julia_count, 1), XP(wedge_julia_power, 1),
   XP(wedge_julia_dist, 0), XP(wedge_sph_angle, 0), XP(wedge_sph_hole, 0), XP(wedge_sph_count, 1),
   XP(wedge_sph_swirl, 0), XP(whorl_inside, 0), XP(whorl_outside, 0), XP(waves2_scalex, 0),
   XP(waves2_scaley, 0), XP(waves2_freqx, 0), XP(waves2_freqy, 0), XP(auger_freq, 1), XP(auger_weight, 0.5),
   XP(auger_sym, 0), XP(auger_scale, 1), XP(flux_spread, 0), XP(mobius_re_a, 0), XP(mobius_re_b, 0),
   XP(mobius_re_c, 0), XP(mobius_re_d, 0), XP(mobius_im_a, 0), XP(mobius_im_b, 0), XP(mobius_im_c, 0),
   XP(mobius_im_d, 0),
};
/* other spellings some writers use (parser.c:1141-1156, 1229-1243) */
static const struct {
   const char *alias, *name;
} k_xform_aliases[] = {
   {"oscilloscope_separation", "oscope_separation"}, {"oscilloscope_frequency", "oscope_frequency"},
   {"oscilloscope_amplitude", "oscope_amplitude"},   {"oscilloscope_damping", "oscope_damping"},
   {"Re_A", "mobius_re_a"}, {"Re_B", "mobius_re_b"}, {"Re_C", "mobius_re_c"}, {"Re_D", "mobius_re_d"},
   {"Im_A", "mobius_im_a"}, {"Im_B", "mobius_im_b"}, {"Im_C", "mobius_im_c"}, {"Im_D", "mobius_im_d"},
};
/* derived constants of a fresh xform that the precalc step overwrites anyway */
static void xform_derived_defaults(flam3_xform *x) {
   x->julian_rN = 1.0;
   x->julian_cn = 0.5;
   x->juliascope_rN = 1.0;
   x->juliascope_cn = 0.5;
   x->radialBlur_zoomvar = 1.0;
   x->wedgeJulia_cn = 0.5;
   x->wedgeJulia_rN = 1.0;
}

/* <flame> attributes that are one number: name, kind (d double, i int), field */
#define GA(attr, kind, field) {attr, kind, offsetof(flam3_genome, field)}
static const struct {
   const char *name;
   char kind;
   size_t off;
} k_flame_numbers[] = {
   GA("time", 'd', time), GA("hsv_rgb_palette_blend", 'd', hsv_rgb_palette_blend), GA("palette", 'i', palette_index),
   GA("scale", 'd', pixels_per_unit), GA("rotate", 'd', rotate), GA("zoom", 'd', zoom),
   GA("oversample", 'i', spatial_oversample), GA("supersample", 'i', spatial_oversample),
   GA("filter", 'd', spatial_filter_radius), GA("temporal_filter_width", 'd', temporal_filter_width),
   GA("temporal_filter_exp", 'd', temporal_filter_exp), GA("quality", 'd', sample_density),
   GA("passes", 'i', nbatches), GA("temporal_samples", 'i', ntemporal_samples), GA("brightness", 'd', brightness),
   GA("gamma", 'd', gamma), GA("vibrancy", 'd', vibrancy), GA("estimator_radius", 'd', estimator),
   GA("estimator_minimum", 'd', estimator_minimum), GA("estimator_curve", 'd', estimator_curve),
   GA("gamma_threshold", 'd', gam_lin_thresh),
};
/* <flame> attributes that are one word of a fixed set */
typedef struct {
   const char *word;
   int value;
} f3b_word;
static const f3b_word k_filter_shapes[] = {
   {"gaussian", flam3_gaussian_kernel}, {"hermite", flam3_hermite_kernel}, {"box", flam3_box_kernel},
   {"triangle", flam3_triangle_kernel}, {"bell", flam3_bell_kernel}, {"bspline", flam3_b_spline_kernel},
   {"mitchell", flam3_mitchell_kernel}, {"blackman", flam3_blackman_kernel}, {"catrom", flam3_catrom_kernel},
   {"hanning", flam3_hanning_kernel}, {"hamming", flam3_hamming_kernel}, {"lanczos3", flam3_lanczos3_kernel},
   {"lanczos2", flam3_lanczos2_kernel}, {"quadratic", flam3_quadratic_kernel}, {NULL, 0}};
static const f3b_word k_interpolation[] = {{"linear", flam3_interpolation_linear}, {"smooth", flam3_interpolation_smooth}, {NULL, 0}};
static const f3b_word k_palette_interp[] = {{"hsv", flam3_palette_interpolation_hsv}, {"sweep", flam3_palette_interpolation_sweep},
                                            {"hsv_circular", flam3_palette_interpolation_hsv_circular},
                                            {"rgb", flam3_palette_interpolation_rgb}, {NULL, 0}};
static const f3b_word k_interp_type[] = {{"linear", flam3_inttype_linear}, {"log", flam3_inttype_log},
                                         {"old", flam3_inttype_compat}, {"older", flam3_inttype_older}, {NULL, 0}};
static const f3b_word k_temporal[] = {{"box", flam3_temporal_box}, {"gaussian", flam3_temporal_gaussian},
                                      {"exp", flam3_temporal_exp}, {NULL, 0}};
static const f3b_word k_palette_mode[] = {{"step", flam3_palette_mode_step}, {"linear", flam3_palette_mode_linear}, {NULL, 0}};
static const struct {
   const char *name;
   const f3b_word *words;
   size_t off;
   const char *what;
} k_flame_words[] = {
   {"filter_shape", k_filter_shapes, offsetof(flam3_genome, spatial_filter_select), "kernel shape"},
   {"interpolation", k_interpolation, offsetof(flam3_genome, interpolation), "interpolation type"},
   {"palette_interpolation", k_palette_interp, offsetof(flam3_genome, palette_interpolation), "palette interpolation type"},
   {"interpolation_space", k_interp_type, offsetof(flam3_genome, interpolation_type), "interpolation_type"},
   {"interpolation_type", k_interp_type, offsetof(flam3_genome, interpolation_type), "interpolation_type"},
   {"temporal_filter_type", k_temporal, offsetof(flam3_genome, temporal_filter_type), "temporal filter"},
   {"palette_mode", k_palette_mode, offsetof(flam3_genome, palette_mode), "palette mode"},
};

/* The reference is built with -ffast-math (Makefile.am:4), which turns every division by a
 * constant into a multiplication by its rounded reciprocal. Colours are n / 255 all over the
 * parser and n / 255 != n * (1 / 255) in the last bit for some n - enough to move a palette
 * entry across a truncation boundary (132 / 255 * 255 = 132, 132 * (1 / 255) * 255 = 131.99...).
 * The same arithmetic, spelled out, so that the genomes are bit-identical however this file is
 * compiled: */
static const double k_inv255 = 1.0 / 255.0;

/* ---- numbers (flam3_atof / flam3_atoi, parser.c:26-72: the whole string must convert) ---- */
static int g_bad_number;
static double to_double(const char *s) {
   char *end;
   errno = 0;
   const double v = strtod(s, &end);
   if (end != s + strlen(s) || errno) {
      g_bad_number = 1;
      fprintf(stderr, "f3b_parse_flam3: cannot convert '%s' to a number\n", s);
   }
   return v;
}
static int to_int(const char *s) {
   char *end;
   errno = 0;
   const long v = strtol(s, &end, 10);
   if (end != s + strlen(s) || errno) {
      g_bad_number = 1;
      fprintf(stderr, "f3b_parse_flam3: cannot convert '%s' to an integer\n", s);
   }
   return (int)v;
}
/* exactly n numbers and nothing else */
static int to_doubles(const char *s, double *out, int n) {
   const char *p = s;
   for (int i = 0; i < n; i++) {
      char *end;
      errno = 0;
      out[i] = strtod(p, &end);
      if (errno || end == p) return 1;
      p = end;
   }
   while (isspace((unsigned char)*p)) p++;
   return *p != 0;
}

/* ---- genome housekeeping ------------------------------------------------------------------ */
void f3b_genome_init(flam3_genome *cp, int defaults_on) {
   /* clear_cp (flam3.c:1270-1370): the values a genome starts from */
   f3b_genome_free_xforms(cp);
   cp->palette_index = flam3_palette_random;
   cp->center[0] = cp->center[1] = cp->rot_center[0] = cp->rot_center[1] = 0.0;
   cp->gamma = 4.0;
   cp->vibrancy = 1.0;
   cp->contrast = 1.0;
   cp->brightness = 4.0;
   cp->symmetry = 0;
   cp->hue_rotation = 0.0;
   cp->rotate = 0.0;
   cp->pixels_per_unit = 50;
   cp->interpolation = flam3_interpolation_linear;
   cp->palette_interpolation = flam3_palette_interpolation_hsv_circular;
   cp->hsv_rgb_palette_blend = 0.0;
   cp->genome_index = 0;
   memset(cp->parent_fname, 0, flam3_parent_fn_len);
   cp->highlight_power = -1.0;
   if (defaults_on) {
      cp->background[0] = cp->background[1] = cp->background[2] = 0.0;
      cp->width = cp->height = 100;
      cp->spatial_oversample = 1;
      cp->spatial_filter_radius = 0.5;
      cp->zoom = 0.0;
      cp->sample_density = 1;
      cp->estimator = 9.0;
      cp->estimator_minimum = 0.0;
      cp->estimator_curve = 0.4;
      cp->gam_lin_thresh = 0.01;
      cp->nbatches = 1;
      cp->ntemporal_samples = 1000;
      cp->spatial_filter_select = flam3_gaussian_kernel;
      cp->interpolation_type = flam3_inttype_log;
      cp->temporal_filter_type = flam3_temporal_box;
      cp->temporal_filter_width = 1.0;
      cp->temporal_filter_exp = 0.0;
      cp->palette_mode = flam3_palette_mode_step;
   } else {
      cp->background[0] = cp->background[1] = cp->background[2] = -1.0;
      cp->zoom = 999999999;
      cp->spatial_oversample = -1;
      cp->spatial_filter_radius = -1;
      cp->width = cp->height = -1;
      cp->sample_density = -1;
      cp->estimator = cp->estimator_minimum = cp->estimator_curve = cp->gam_lin_thresh = -1;
      cp->nbatches = 0;
      cp->ntemporal_samples = 0;
      cp->spatial_filter_select = -1;
      cp->interpolation_type = -1;
      cp->temporal_filter_type = -1;
      cp->temporal_filter_width = -1;
      cp->temporal_filter_exp = -999;
      cp->palette_mode = -1;
   }
   cp->final_xform_enable = 0;
   cp->final_xform_index = -1;
}

void f3b_genome_free_xforms(flam3_genome *cp) {
   if (cp->xform && cp->num_xforms > 0) {
      const int nstd = cp->num_xforms - (cp->final_xform_index >= 0);
      if (cp->chaos) {
         for (int i = 0; i < nstd; i++) free(cp->chaos[i]);
         free(cp->chaos);
      }
      for (int i = 0; i < cp->num_xforms; i++)
         if (cp->xform[i].num_motion > 0) free(cp->xform[i].motion);
      free(cp->xform);
   }
   cp->xform = NULL;
   cp->chaos = NULL;
   cp->num_xforms = 0;
   cp->final_xform_index = -1;
}

static void xform_defaults(flam3_xform *x, int position) {
   memset(x, 0, sizeof(*x));
   x->color_speed = 0.5;
   x->animate = 1.0;
   x->color = position & 1;
   x->opacity = 1.0;
   x->var[0] = 1.0;
   x->c[0][0] = x->c[1][1] = 1.0;
   x->post[0][0] = x->post[1][1] = 1.0;
   for (size_t i = 0; i < sizeof(k_xform_params) / sizeof(k_xform_params[0]); i++)
      *(double *)((char *)x + k_xform_params[i].off) = k_xform_params[i].dflt;
   xform_derived_defaults(x);
}

/* one more xform: a standard one goes in front of the final xform (which stays last) and
 * widens the chaos matrix with ones; the final xform has no row or column in it */
int f3b_genome_add_xform(flam3_genome *cp, int is_final) {
   const int had_final = cp->final_xform_index >= 0;
   if (is_final && had_final) return -1;
   const int old_std = cp->num_xforms - had_final;
   flam3_xform *nx = (flam3_xform *)realloc(cp->xform, (size_t)(cp->num_xforms + 1) * sizeof(flam3_xform));
   if (!nx) return -1;
   cp->xform = nx;
   int at = cp->num_xforms;
   if (!is_final && had_final) { /* keep the final xform at the end */
      cp->xform[cp->num_xforms] = cp->xform[cp->final_xform_index];
      at = cp->final_xform_index;
      cp->final_xform_index = cp->num_xforms;
   }
   cp->num_xforms++;
   xform_defaults(&cp->xform[at], at);
   if (is_final) {
      cp->final_xform_enable = 1;
      cp->final_xform_index = at;
      return at;
   }
   const int nstd = old_std + 1;
   double **rows = (double **)realloc(cp->chaos, (size_t)nstd * sizeof(double *));
   if (!rows) return -1;
   cp->chaos = rows;
   cp->chaos[old_std] = NULL;
   for (int i = 0; i < nstd; i++) {
      double *r = (double *)realloc(cp->chaos[i], (size_t)nstd * sizeof(double));
      if (!r) return -1;
      cp->chaos[i] = r;
      for (int j = (i < old_std ? old_std : 0); j < nstd; j++) r[j] = 1.0;
   }
   return at;
}

int f3b_genome_copy(flam3_genome *dest, const flam3_genome *src) {
   f3b_genome_free_xforms(dest);
   memcpy(dest, src, sizeof(*dest));
   dest->xform = NULL;
   dest->chaos = NULL;
   dest->num_xforms = 0;
   dest->final_xform_index = -1;
   const int nstd = src->num_xforms - (src->final_xform_index >= 0);
   for (int i = 0; i < src->num_xforms; i++) {
      const int is_final = (i == src->final_xform_index);
      const int at = f3b_genome_add_xform(dest, is_final);
      if (at < 0) return 1;
      dest->xform[at] = src->xform[i];
      dest->xform[at].motion = NULL;
      dest->xform[at].num_motion = 0;
      if (src->xform[i].num_motion > 0) {
         const size_t bytes = (size_t)src->xform[i].num_motion * sizeof(flam3_xform);
         dest->xform[at].motion = (struct xform *)malloc(bytes);
         if (!dest->xform[at].motion) return 1;
         memcpy(dest->xform[at].motion, src->xform[i].motion, bytes);
         dest->xform[at].num_motion = src->xform[i].num_motion;
      }
   }
   dest->final_xform_enable = src->final_xform_enable;
   for (int i = 0; i < nstd; i++) memcpy(dest->chaos[i], src->chaos[i], (size_t)nstd * sizeof(double));
   return 0;
}

/* ---- <symmetry kind="n"/> (flam3_add_symmetry, flam3.c:2558-2641) ----------------------- */
/* n >= 2: n-1 rotations by multiples of 2 pi / n; n <= -1: a reflection first, then the
 * rotations of |n|. The added xforms have weight 1, colour speed 0, do not animate, and are
 * put in order of their determinant (all colour speeds are 0; interpolation.c:117-147) by a
 * stable sort - over the LAST `added` entries of the xform array, as the reference does. */
static double round_6(double x) {
   x *= 1e6;
   if (x < 0) x -= 1.0;
   return 1e-6 * (int)(x + 0.5);
}
static int symmetry_order(const flam3_xform *a, const flam3_xform *b) {
   double ad = a->c[0][0] * a->c[1][1] - a->c[0][1] * a->c[1][0];
   double bd = b->c[0][0] * b->c[1][1] - b->c[0][1] * b->c[1][0];
   if (a->color_speed > b->color_speed) return 1;
   if (a->color_speed < b->color_speed) return -1;
   if (a->color_speed) {
      if (ad < 0) return -1;
      if (bd < 0) return 1;
      ad = atan2(a->c[0][0], a->c[0][1]);
      bd = atan2(b->c[0][0], b->c[0][1]);
   }
   return ad < bd ? -1 : (ad > bd ? 1 : 0);
}
static int add_symmetry(flam3_genome *cp, int kind) {
   if (kind == 0) {
      /* a symmetry drawn from the C library's random(), with the reference's odds and in its
       * order of draws (flam3.c:2564-2579): the same state of random() gives the same genome */
      static const int common[] = {-4, -3, -2, -2, -2, -1, -1, -1, 2, 2, 2, 3, 3, 4, 4};
      if (random() & 1) kind = common[random() % (long)(sizeof(common) / sizeof(common[0]))];
      else if (random() & 31) kind = (int)(random() % 13) - 6;
      else kind = (int)(random() % 51) - 25;
   }
   if (kind == 1 || kind == 0) return 0;
   cp->symmetry = kind;
   int added = 0;
   for (int pass = (kind < 0 ? 0 : 1), n = kind < 0 ? -kind : kind, k = 1; pass < 2 || k < n;) {
      const int at = f3b_genome_add_xform(cp, 0);
      if (at < 0) return -1;
      flam3_xform *x = &cp->xform[at];
      x->density = 1.0;
      x->color_speed = 0.0;
      x->animate = 0.0;
      memset(x->var, 0, sizeof(x->var));
      x->var[0] = 1.0;
      if (pass == 0) { /* the reflection of a negative kind */
         x->color = 1.0;
         x->c[0][0] = -1.0;
         x->c[0][1] = x->c[1][0] = 0.0;
         x->c[1][1] = 1.0;
         pass = 2;
      } else {
         pass = 2;
         const double a = 2 * M_PI / n;
         x->color = n < 3 ? 0.0 : (k - 1.0) / (n - 2.0);
         x->c[0][0] = round_6(cos(k * a));
         x->c[0][1] = round_6(sin(k * a));
         x->c[1][0] = round_6(-x->c[0][1]);
         x->c[1][1] = x->c[0][0];
         k++;
      }
      x->c[2][0] = x->c[2][1] = 0.0;
      added++;
   }
   /* stable insertion sort of the last `added` entries */
   flam3_xform *base = cp->xform + (cp->num_xforms - added);
   for (int i = 1; i < added; i++) {
      flam3_xform key = base[i];
      int j = i - 1;
      while (j >= 0 && symmetry_order(&base[j], &key) > 0) {
         base[j + 1] = base[j];
         j--;
      }
      base[j + 1] = key;
   }
   return added;
}

/* ---- colours ---------------------------------------------------------------------------- */
/* `count` colours of 2 hex digits per channel (parser.c:83-133): chan 3 = RRGGBB, 4 = RRGGBBAA,
 * -4 = 00RRGGBB; whitespace between colours */
static int hex_colors(const char *s, flam3_genome *cp, int count, int chan) {
   const int digits = 2 * (chan < 0 ? -chan : chan);
   size_t at = 0;
   while (isspace((unsigned char)s[at])) at++;
   int n = 0;
   do {
      unsigned v[4] = {0, 0, 0, 255};
      const int lead = chan == -4 ? 2 : 0, nch = chan == 4 ? 4 : 3;
      if (lead && (s[at] != '0' || s[at + 1] != '0')) goto bad;
      for (int c = 0; c < nch; c++) {
         char two[3] = {s[at + lead + 2 * c], s[at + lead + 2 * c + 1], 0};
         if (!isxdigit((unsigned char)two[0]) || !isxdigit((unsigned char)two[1])) goto bad;
         v[c] = (unsigned)strtoul(two, NULL, 16);
      }
      at += (size_t)digits;
      while (isspace((unsigned char)s[at])) at++;
      if (n < 256) {
         for (int c = 0; c < 4; c++) cp->palette[n].color[c] = v[c] * k_inv255;
         cp->palette[n].index = n;
      }
      n++;
   } while (n < count);
   if (s[at]) {
      fprintf(stderr, "f3b_parse_flam3: extra data at the end of hex colour data '%.20s'\n", s + at);
      return 1;
   }
   return 0;
bad:
   fprintf(stderr, "f3b_parse_flam3: problem reading hexadecimal colour data\n");
   return 1;
}

/* palette entries that were not given are interpolated between their neighbours, wrapping
 * around (parser.c:135-218) */
static void fill_missing_colors(flam3_genome *cp) {
   int first = -1, last = -1;
   for (int i = 0; i < 256; i++)
      if (cp->palette[i].index >= 0) {
         if (first < 0) first = i;
         last = i;
      }
   if (first < 0) {
      for (int i = 0; i < 256; i++) cp->palette[i].index = i;
      return;
   }
   for (int i = 0; i < 256;) {
      if (cp->palette[i].index >= 0) {
         i++;
         continue;
      }
      int end = i;
      while (end + 1 < 256 && cp->palette[end + 1].index < 0) end++;
      /* neighbours of the gap [i, end]: positions for the weights, entries for the colours */
      int lpos = i - 1, rpos = end + 1, lcol = lpos, rcol = rpos;
      if (lpos < 0) {
         lpos = last - 256;
         lcol = last;
      }
      if (rpos > 255) {
         rpos = first + 256;
         rcol = first;
      }
      /* one reciprocal per gap, as the reference's -ffast-math build hoists it (see k_inv255) */
      const double inv_span = 1.0 / (double)(rpos - lpos);
      for (int j = i; j <= end; j++) {
         const double t = (j - lpos) * inv_span;
         for (int k = 0; k < 4; k++)
            cp->palette[j].color[k] = cp->palette[lcol].color[k] * (1.0 - t) + cp->palette[rcol].color[k] * t;
         cp->palette[j].index = j;
      }
      i = end + 1;
   }
}

/* hexcone model, h in [0,6) (palettes.c:199-272). The hue is the difference of two of
 * (max - channel) / del; the reference's -ffast-math build folds that into ONE division of the
 * channel difference, (g - b) / del and so on (see k_inv255) - the same here, to the bit. */
static void rgb_to_hsv(const double *rgb, double *hsv) {
   const double r = rgb[0], g = rgb[1], b = rgb[2];
   const double mx = r >= g ? (r >= b ? r : b) : (g >= b ? g : b);
   const double mn = r <= g ? (r <= b ? r : b) : (g <= b ? g : b);
   const double del = mx - mn;
   const double s = mx != 0.0 ? del / mx : 0.0;
   double h = 0.0;
   if (s != 0.0) {
      if (r == mx) h = (g - b) / del;
      else if (g == mx) h = (b - r) / del + 2.0;
      else if (b == mx) h = (r - g) / del + 4.0;
      if (h < 0) h += 6;
   }
   hsv[0] = h;
   hsv[1] = s;
   hsv[2] = mx;
}
static void hsv_to_rgb(const double *hsv, double *rgb) {
   double h = hsv[0];
   const double s = hsv[1], v = hsv[2];
   while (h >= 6.0) h -= 6.0;
   while (h < 0.0) h += 6.0;
   const int j = (int)floor(h);
   const double f = h - j, p = v * (1 - s), q = v * (1 - s * f), t = v * (1 - s * (1 - f));
   const double tab[6][3] = {{v, t, p}, {q, v, p}, {p, v, t}, {p, q, v}, {t, p, v}, {v, p, q}};
   memcpy(rgb, tab[j % 6], 3 * sizeof(double));
}

/* the palette library: <palette number="N" name="..." data="00RRGGBB x 256"/> */
static struct {
   int number;
   unsigned char rgb[256][3];
} *g_palettes;
static int g_npalettes = -1; /* -1: not read yet */

static void collect_palettes(xmlNode *n) {
   for (; n; n = n->next) {
      if (n->type == XML_ELEMENT_NODE && !xmlStrcmp(n->name, (const xmlChar *)"palette")) {
         char *num = (char *)xmlGetProp(n, (const xmlChar *)"number");
         char *data = (char *)xmlGetProp(n, (const xmlChar *)"data");
         if (data) {
            void *grown = realloc(g_palettes, (size_t)(g_npalettes + 1) * sizeof(*g_palettes));
            if (grown) {
               g_palettes = grown;
               g_palettes[g_npalettes].number = num ? atoi(num) : 0;
               size_t at = 0;
               int ok = 1;
               for (int c = 0; c < 256 && ok; c++) {
                  unsigned r, g, b;
                  if (sscanf(data + at, "00%2x%2x%2x", &r, &g, &b) != 3) ok = 0;
                  at += 8;
                  while (ok && isspace((unsigned char)data[at])) at++;
                  g_palettes[g_npalettes].rgb[c][0] = (unsigned char)r;
                  g_palettes[g_npalettes].rgb[c][1] = (unsigned char)g;
                  g_palettes[g_npalettes].rgb[c][2] = (unsigned char)b;
               }
               if (ok) g_npalettes++;
               else fprintf(stderr, "f3b_get_palette: problem reading hexadecimal colour data of palette %s\n", num ? num : "?");
            }
         }
         if (num) xmlFree(num);
         if (data) xmlFree(data);
      } else
         collect_palettes(n->children);
   }
}

static char *read_file(const char *path) {
   FILE *f = fopen(path, "rb");
   if (!f) return NULL;
   size_t cap = 1 << 16, n = 0;
   char *s = (char *)malloc(cap);
   for (size_t got; s && (got = fread(s + n, 1, cap - n - 1, f)) > 0;) {
      n += got;
      if (n + 1 >= cap) s = (char *)realloc(s, cap *= 2);
   }
   fclose(f);
   if (s) s[n] = 0;
   return s;
}

int f3b_get_palette(int number, flam3_palette pal, double hue_rotation) {
   for (int i = 0; i < 256; i++) { /* all white in case of problems, like the reference */
      pal[i].index = i;
      for (int j = 0; j < 4; j++) pal[i].color[j] = 1.0;
   }
   if (g_npalettes < 0) {
      const char *path = getenv("flam3_palettes");
      if (!path) path = "flam3-palettes.xml";
      char *text = read_file(path);
      if (!text) {
         fprintf(stderr, "f3b_get_palette: could not open palette file %s (set flam3_palettes)\n", path);
         return -1;
      }
      xmlDocPtr doc = xmlReadMemory(text, (int)strlen(text), path, NULL, XML_PARSE_NONET);
      free(text);
      if (!doc) {
         fprintf(stderr, "f3b_get_palette: error parsing %s\n", path);
         return -1;
      }
      g_npalettes = 0;
      collect_palettes(xmlDocGetRootElement(doc));
      xmlFreeDoc(doc);
   }
   if (g_npalettes == 0) return -1;
   if (number == flam3_palette_random) number = g_palettes[random() % g_npalettes].number;
   for (int p = 0; p < g_npalettes; p++) {
      if (g_palettes[p].number != number) continue;
      for (int i = 0; i < 256; i++) {
         double rgb[3], hsv[3];
         for (int j = 0; j < 3; j++) rgb[j] = g_palettes[p].rgb[i][j] * k_inv255;
         rgb_to_hsv(rgb, hsv);
         hsv[0] += hue_rotation * 6.0;
         hsv_to_rgb(hsv, rgb);
         for (int j = 0; j < 3; j++) pal[i].color[j] = rgb[j];
         pal[i].color[3] = 1.0;
      }
      return number;
   }
   fprintf(stderr, "warning: palette number %d not found, using white.\n", number);
   return -1;
}

/* ---- <xform> / <finalxform> ------------------------------------------------------------- */
typedef struct {
   int from, to;
   double scalar;
} xaos_entry;

static int xform_from_xml(xmlNode *node, flam3_xform *x, int std_index, xaos_entry **xaos, int *nxaos, int motion) {
   if (!node->properties) {
      fprintf(stderr, "Error: No attributes for element.\n");
      return 1;
   }
   for (xmlAttrPtr a = node->properties; a; a = a->next) {
      const char *name = (const char *)a->name;
      char *val = (char *)xmlGetProp(node, a->name);
      int rc = 0;
      if (!strcmp(name, "weight")) x->density = to_double(val);
      else if (!strcmp(name, "symmetry")) { /* deprecated spelling of color_speed + animate */
         const double s = to_double(val);
         x->color_speed = (1.0 - s) / 2.0;
         x->animate = s > 0 ? 0 : 1;
      } else if (!strcmp(name, "color_speed")) x->color_speed = to_double(val);
      else if (!strcmp(name, "animate")) x->animate = to_double(val);
      else if (!strcmp(name, "opacity")) x->opacity = to_double(val);
      else if (!strcmp(name, "plotmode")) {
         if (motion) {
            fprintf(stderr, "error: motion element cannot have a plotmode attribute.\n");
            rc = 1;
         } else if (!strcmp(val, "off")) x->opacity = 0.0;
      } else if (!strcmp(name, "color")) { /* one coordinate; an old second one is read and dropped */
         double two[2];
         x->color = 0.0;
         if (to_doubles(val, two, 2) == 0) x->color = two[0];
         else if (to_doubles(val, two, 1) == 0) x->color = two[0];
         else {
            fprintf(stderr, "Error: malformed xform color attribute '%s'\n", val);
            rc = 1;
         }
      } else if (!strcmp(name, "coefs") || !strcmp(name, "post")) {
         double six[6];
         if (to_doubles(val, six, 6)) {
            fprintf(stderr, "error: bad value in %s attribute '%s'\n", name, val);
            rc = 1;
         } else {
            double(*m)[2] = name[0] == 'c' ? x->c : x->post;
            for (int k = 0; k < 3; k++)
               for (int j = 0; j < 2; j++) m[k][j] = six[2 * k + j];
         }
      } else if (!strcmp(name, "var1")) {
         const int v = to_int(val);
         if (v < 0 || v >= flam3_nvariations) {
            fprintf(stderr, "Error:  Bad variation (%d)\n", v);
            rc = 1;
         } else {
            memset(x->var, 0, sizeof(x->var));
            x->var[v] = 1.0;
         }
      } else if (!strcmp(name, "var")) {
         if (to_doubles(val, x->var, flam3_nvariations)) {
            fprintf(stderr, "error: bad value in var attribute '%s'\n", val);
            rc = 1;
         }
      } else if (!strcmp(name, "chaos") && motion) {
         fprintf(stderr, "error: motion element cannot have a chaos attribute.\n");
         rc = 1;
      } else if (!strcmp(name, "chaos")) {
         /* entries that are not 1 are remembered and applied when all xforms are known */
         int to = 0;
         for (char *tok = strtok(val, " "); tok && !rc; tok = strtok(NULL, " "), to++) {
            const double s = to_double(tok);
            if (s == 1.0) continue;
            xaos_entry *grown = (xaos_entry *)realloc(*xaos, (size_t)(*nxaos + 1) * sizeof(xaos_entry));
            if (!grown) rc = 1;
            else {
               *xaos = grown;
               (*xaos)[*nxaos].from = std_index;
               (*xaos)[*nxaos].to = to;
               (*xaos)[*nxaos].scalar = s;
               (*nxaos)++;
            }
         }
      } else if (!strcmp(name, "motion_frequency")) x->motion_freq = to_int(val);
      else if (!strcmp(name, "motion_function")) {
         if (!strcmp(val, "sin")) x->motion_func = MOTION_SIN;
         else if (!strcmp(val, "triangle")) x->motion_func = MOTION_TRIANGLE;
         else if (!strcmp(val, "hill")) x->motion_func = MOTION_HILL;
         else {
            fprintf(stderr, "Error: unknown motion function '%s'\n", val);
            rc = 1;
         }
      } else if (!strcmp(name, "super_shape_rnd")) {
         const double v = to_double(val);
         x->super_shape_rnd = v < 0 ? 0 : (v > 1 ? 1 : v);
      } else {
         const char *field = name;
         for (size_t i = 0; i < sizeof(k_xform_aliases) / sizeof(k_xform_aliases[0]); i++)
            if (!strcmp(name, k_xform_aliases[i].alias)) field = k_xform_aliases[i].name;
         int done = 0;
         for (size_t i = 0; i < sizeof(k_xform_params) / sizeof(k_xform_params[0]) && !done; i++)
            if (!strcmp(field, k_xform_params[i].name)) {
               *(double *)((char *)x + k_xform_params[i].off) = to_double(val);
               done = 1;
            }
         for (int v = 0; v < flam3_nvariations && !done; v++)
            if (!strcmp(name, k_variation_names[v])) {
               x->var[v] = to_double(val);
               done = 1;
            }
         if (!done) fprintf(stderr, "Warning: unrecognized variation %s.  Ignoring.\n", name);
      }
      xmlFree(val);
      if (rc) return 1;
   }
   if (motion) return 0;
   if (x->motion_freq != 0 || x->motion_func != 0) {
      fprintf(stderr, "Error: Motion parameters should not be specified in xforms.\n");
      return 1;
   }
   /* <motion> children (parser.c:763-777): xforms that start from all zeros and say what to add
    * to their parent, as a function of time, when flam3-genome builds sequences; a render does
    * not look at them */
   for (xmlNode *c = node->children; c; c = c->next)
      if (c->type == XML_ELEMENT_NODE && !xmlStrcmp(c->name, (const xmlChar *)"motion")) {
         flam3_xform *grown = (flam3_xform *)realloc(x->motion, (size_t)(x->num_motion + 1) * sizeof(flam3_xform));
         if (!grown) return 1;
         x->motion = grown;
         memset(&x->motion[x->num_motion], 0, sizeof(flam3_xform));
         x->num_motion++;
         if (xform_from_xml(c, &x->motion[x->num_motion - 1], 0, NULL, NULL, 1)) return 1;
      }
   return 0;
}

/* ---- <flame> ---------------------------------------------------------------------------- */
static int flame_from_xml(xmlNode *flame, flam3_genome *cp) {
   if (!flame->properties) {
      fprintf(stderr, "Error : <flame> element has no attributes.\n");
      return 1;
   }
   for (int i = 0; i < 256; i++) {
      memset(cp->palette[i].color, 0, sizeof(cp->palette[i].color));
      cp->palette[i].index = -1;
   }
   memset(cp->flame_name, 0, flam3_name_len + 1);
   g_bad_number = 0;
   int solo = -1;

   for (xmlAttrPtr a = flame->properties; a; a = a->next) {
      const char *name = (const char *)a->name;
      char *val = (char *)xmlGetProp(flame, a->name);
      int rc = 0, done = 0;
      for (size_t i = 0; i < sizeof(k_flame_numbers) / sizeof(k_flame_numbers[0]) && !done; i++)
         if (!strcmp(name, k_flame_numbers[i].name)) {
            if (k_flame_numbers[i].kind == 'd') *(double *)((char *)cp + k_flame_numbers[i].off) = to_double(val);
            else *(int *)((char *)cp + k_flame_numbers[i].off) = to_int(val);
            done = 1;
         }
      for (size_t i = 0; i < sizeof(k_flame_words) / sizeof(k_flame_words[0]) && !done; i++)
         if (!strcmp(name, k_flame_words[i].name)) {
            const f3b_word *w = k_flame_words[i].words;
            while (w->word && strcmp(w->word, val)) w++;
            if (w->word) *(int *)((char *)cp + k_flame_words[i].off) = w->value;
            else fprintf(stderr, "warning: unrecognized %s %s.\n", k_flame_words[i].what, val);
            done = 1;
         }
      if (done) {
      } else if (!strcmp(name, "name")) {
         strncpy(cp->flame_name, val, flam3_name_len);
         for (int i = (int)strlen(cp->flame_name) - 2; i >= 0; i--) /* as the reference: not the last one */
            if (isspace((unsigned char)cp->flame_name[i])) cp->flame_name[i] = '_';
      } else if (!strcmp(name, "size")) {
         double wh[2];
         if (to_doubles(val, wh, 2) || wh[0] != (int)wh[0] || wh[1] != (int)wh[1]) {
            fprintf(stderr, "error: invalid size attribute '%s'\n", val);
            rc = 1;
         } else {
            cp->width = (int)wh[0];
            cp->height = (int)wh[1];
         }
      } else if (!strcmp(name, "center")) {
         if (to_doubles(val, cp->center, 2)) {
            fprintf(stderr, "error: invalid center attribute '%s'\n", val);
            rc = 1;
         } else {
            cp->rot_center[0] = cp->center[0];
            cp->rot_center[1] = cp->center[1];
         }
      } else if (!strcmp(name, "background")) {
         if (to_doubles(val, cp->background, 3)) {
            fprintf(stderr, "error: invalid background attribute '%s'\n", val);
            rc = 1;
         }
      } else if (!strcmp(name, "highlight_power")) cp->highlight_power = atof(val);
      else if (!strcmp(name, "hue")) cp->hue_rotation = fmod(to_double(val), 1.0);
      else if (!strcmp(name, "soloxform")) solo = (int)to_double(val);
      /* anything else (version, brightness aliases of other programs, ...) is ignored */
      xmlFree(val);
      if (rc) return 1;
   }

   xaos_entry *xaos = NULL;
   int nxaos = 0, nstd = 0, rc = 0;
   for (xmlNode *c = flame->children; c && !rc; c = c->next) {
      if (c->type != XML_ELEMENT_NODE) continue;
      const char *tag = (const char *)c->name;
      if (!strcmp(tag, "color")) {
         int index = -1;
         double rgba[4] = {0, 0, 0, 255.0};
         for (xmlAttrPtr a = c->properties; a && !rc; a = a->next) {
            char *val = (char *)xmlGetProp(c, a->name);
            const char *an = (const char *)a->name;
            if (!strcmp(an, "index")) index = to_int(val);
            else if (!strcmp(an, "rgb")) rc = to_doubles(val, rgba, 3);
            else if (!strcmp(an, "rgba")) rc = to_doubles(val, rgba, 4);
            else if (!strcmp(an, "a")) rc = to_doubles(val, rgba + 3, 1);
            else {
               fprintf(stderr, "Error:  Unknown color attribute '%s'\n", an);
               rc = 1;
            }
            if (rc) fprintf(stderr, "error: invalid colour attribute '%s'\n", val);
            xmlFree(val);
         }
         if (!rc && (index < 0 || index > 255)) {
            fprintf(stderr, "Error:  Color element with bad/missing index attribute (%d)\n", index);
            rc = 1;
         }
         if (!rc) { /* premultiplied by alpha */
            const double alpha = rgba[3] * k_inv255;
            cp->palette[index].color[3] = alpha;
            /* (r / 255) * alpha, in this order: what the reference's build makes of alpha * r / 255 */
            for (int k = 0; k < 3; k++) cp->palette[index].color[k] = (rgba[k] * k_inv255) * alpha;
            cp->palette[index].index = index;
         }
      } else if (!strcmp(tag, "colors")) {
         int count = 0;
         for (xmlAttrPtr a = c->properties; a && !rc; a = a->next) {
            char *val = (char *)xmlGetProp(c, a->name);
            if (!strcmp((const char *)a->name, "count")) count = to_int(val);
            else if (!strcmp((const char *)a->name, "data")) rc = hex_colors(val, cp, count, -4);
            else {
               fprintf(stderr, "Error:  Unknown color attribute '%s'\n", (const char *)a->name);
               rc = 1;
            }
            xmlFree(val);
         }
      } else if (!strcmp(tag, "palette")) {
         int count = 0, bytes = 0, old_style = 0, new_style = 0;
         int index0 = flam3_palette_random, index1 = flam3_palette_random;
         double hue0 = 0.0, hue1 = 0.0, blend = 0.5;
         if (!c->properties) {
            fprintf(stderr, "Error:  No attributes for palette element.\n");
            rc = 1;
         }
         for (xmlAttrPtr a = c->properties; a && !rc; a = a->next) {
            char *val = (char *)xmlGetProp(c, a->name);
            const char *an = (const char *)a->name;
            if (!strcmp(an, "count")) count = to_int(val), new_style++;
            else if (!strcmp(an, "format")) {
               new_style++;
               bytes = !strcmp(val, "RGB") ? 3 : (!strcmp(val, "RGBA") ? 4 : 0);
               if (!bytes) {
                  fprintf(stderr, "Error: Unrecognized palette format string (%s)\n", val);
                  rc = 1;
               }
            } else if (!strcmp(an, "index0")) index0 = to_int(val), old_style++;
            else if (!strcmp(an, "index1")) index1 = to_int(val), old_style++;
            else if (!strcmp(an, "hue0")) hue0 = to_double(val), old_style++;
            else if (!strcmp(an, "hue1")) hue1 = to_double(val), old_style++;
            else if (!strcmp(an, "blend")) blend = to_double(val), old_style++;
            else {
               fprintf(stderr, "Error:  Unknown palette attribute '%s'\n", an);
               rc = 1;
            }
            xmlFree(val);
         }
         if (!rc && old_style && new_style) {
            fprintf(stderr, "Error: mixing of old and new palette tag syntax not allowed.\n");
            rc = 1;
         }
         if (!rc && old_style) {
            /* the pre-2008 way (interpolate_cmap, interpolation.c:149-192): two palettes of the
             * library, rotated in hue, blended in HSV the shorter way round the hue circle */
            flam3_palette p0, p1;
            if (f3b_get_palette(index0, p0, hue0) < 0) fprintf(stderr, "unable to retrieve palette %d, setting to white\n", index0);
            if (f3b_get_palette(index1, p1, hue1) < 0) fprintf(stderr, "unable to retrieve palette %d, setting to white\n", index1);
            for (int i = 0; i < 256; i++) {
               double s5[5], t5[5];
               rgb_to_hsv(p0[i].color, s5);
               rgb_to_hsv(p1[i].color, t5);
               s5[3] = p0[i].color[3];
               t5[3] = p1[i].color[3];
               s5[4] = p0[i].index;
               t5[4] = p1[i].index;
               if (s5[0] - t5[0] > 3.0) s5[0] -= 6.0;
               if (s5[0] - t5[0] < -3.0) s5[0] += 6.0;
               for (int j = 0; j < 5; j++) t5[j] = ((1.0 - blend) * s5[j]) + (blend * t5[j]);
               hsv_to_rgb(t5, cp->palette[i].color);
               cp->palette[i].color[3] = t5[3];
               cp->palette[i].index = t5[4];
            }
         } else if (!rc) {
            char *text = (char *)xmlNodeGetContent(c);
            rc = hex_colors(text ? text : "", cp, count, bytes ? bytes : 3);
            if (text) xmlFree(text);
         }
      } else if (!strcmp(tag, "xform") || !strcmp(tag, "finalxform")) {
         const int is_final = tag[0] == 'f';
         if (is_final && cp->final_xform_index >= 0) {
            fprintf(stderr, "Error:  Cannot specify more than one final xform.\n");
            rc = 1;
            break;
         }
         const int at = f3b_genome_add_xform(cp, is_final);
         if (at < 0) {
            rc = 1;
            break;
         }
         cp->xform[at].var[0] = 0.0;
         rc = xform_from_xml(c, &cp->xform[at], is_final ? -1 : nstd, &xaos, &nxaos, 0);
         if (!rc && is_final && cp->xform[at].density != 0.0) {
            fprintf(stderr, "Error: Final xforms should not have weight specified.\n");
            rc = 1;
         }
         if (!is_final) nstd++;
      } else if (!strcmp(tag, "symmetry")) {
         int kind = 0;
         for (xmlAttrPtr a = c->properties; a && !rc; a = a->next) {
            char *val = (char *)xmlGetProp(c, a->name);
            if (!strcmp((const char *)a->name, "kind")) kind = to_int(val);
            else {
               fprintf(stderr, "Error:  Unknown symmetry attribute '%s'\n", (const char *)a->name);
               rc = 1;
            }
            xmlFree(val);
         }
         if (!rc) {
            const int added = add_symmetry(cp, kind);
            if (added < 0) rc = 1;
            else nstd += added;
         }
      }
      /* <edit> and unknown elements: no effect on rendering */
   }
   if (!rc) {
      for (int i = 0; i < nstd; i++)
         if (solo >= 0 && i != solo) cp->xform[i].opacity = 0.0;
      for (int i = 0; i < nxaos; i++)
         if (xaos[i].from >= 0 && xaos[i].from < nstd && xaos[i].to < nstd) cp->chaos[xaos[i].from][xaos[i].to] = xaos[i].scalar;
      if (g_bad_number) {
         fprintf(stderr, "error: parsing a double or int attribute's value.\n");
         rc = 1;
      }
   }
   free(xaos);
   return rc;
}

static int collect_flames(xmlNode *n, int defaults_on, flam3_genome **all, int *count) {
   for (; n; n = n->next) {
      if (n->type == XML_ELEMENT_NODE && !xmlStrcmp(n->name, (const xmlChar *)"flame")) {
         flam3_genome cp;
         memset(&cp, 0, sizeof(cp));
         f3b_genome_init(&cp, defaults_on);
         if (flame_from_xml(n, &cp)) {
            fprintf(stderr, "error parsing flame element\n");
            f3b_genome_free_xforms(&cp);
            return 1;
         }
         if (cp.palette_index != -1 && f3b_get_palette(cp.palette_index, cp.palette, cp.hue_rotation) < 0)
            fprintf(stderr, "error retrieving palette %d, setting to all white\n", cp.palette_index);
         fill_missing_colors(&cp);
         cp.genome_index = *count;
         flam3_genome *grown = (flam3_genome *)realloc(*all, (size_t)(*count + 1) * sizeof(flam3_genome));
         if (!grown) {
            f3b_genome_free_xforms(&cp);
            return 1;
         }
         *all = grown;
         (*all)[*count] = cp; /* the xform and chaos arrays move with it */
         (*count)++;
      } else if (collect_flames(n->children, defaults_on, all, count))
         return 1;
   }
   return 0;
}

flam3_genome *f3b_parse_flam3(const char *xmltext, int defaults_on, int *ncps) {
   *ncps = 0;
   xmlDocPtr doc = xmlReadMemory(xmltext, (int)strlen(xmltext), "f3b.flam3", NULL, XML_PARSE_NONET);
   if (!doc) {
      fprintf(stderr, "f3b_parse_flam3: failed to parse the XML text\n");
      return NULL;
   }
   flam3_genome *all = NULL;
   int n = 0;
   const int rc = collect_flames(xmlDocGetRootElement(doc), defaults_on, &all, &n);
   xmlFreeDoc(doc);
   if (rc) {
      f3b_free_genomes(all, n);
      return NULL;
   }
   /* what flam3_parse_xml2 does with the whole sequence (flam3.c:1477-1514): no smooth
    * interpolation on the first and last segment, consecutive `rotate` values brought within
    * 180 degrees of each other (not for the two oldest interpolation types) */
   if (n >= 1 && all[0].interpolation == flam3_interpolation_smooth) {
      fprintf(stderr, "Warning: smooth interpolation cannot be used for first segment.\n         switching to linear.\n");
      all[0].interpolation = flam3_interpolation_linear;
   }
   if (n >= 2 && all[n - 2].interpolation == flam3_interpolation_smooth) {
      fprintf(stderr, "Warning: smooth interpolation cannot be used for last segment.\n         switching to linear.\n");
      all[n - 2].interpolation = flam3_interpolation_linear;
   }
   for (int i = 1; i < n; i++) {
      if (all[i - 1].interpolation_type == flam3_inttype_compat || all[i - 1].interpolation_type == flam3_inttype_older)
         continue;
      while (all[i].rotate < all[i - 1].rotate - 180) all[i].rotate += 360;
      while (all[i].rotate > all[i - 1].rotate + 180) all[i].rotate -= 360;
   }
   *ncps = n;
   return all;
}

void f3b_free_genomes(flam3_genome *g, int n) {
   if (!g) return;
   for (int i = 0; i < n; i++) f3b_genome_free_xforms(&g[i]);
   free(g);
}


/* ---- between key frames (flam3.c:717-802, interpolation.c) ------------------------------ */
/* What flam3_interpolate gives flam3_render for a time between key frames: the key frames are
 * padded to the same xforms (flam3_align), everything is blended with the segment's weights -
 * two of them, or four Catmull-Rom weights on smooth segments - the affine parts in log-polar
 * form; with stagger the xforms take their turns. Motion elements are not read by this front
 * end, so there are none to apply. Pinned
 * against the reference byte for byte (tests/test_cpu_frontend.py). */

/* which variation a parameter belongs to: its name is the variation's name + "_" + something;
 * the longest such variation wins (wedge_julia_angle is wedge_julia's, not wedge's) */
static int param_variation(const char *pname) {
   int best = -1;
   size_t best_len = 0;
   for (int v = 0; v < flam3_nvariations; v++) {
      const char *vn = v == VAR_OSCILLOSCOPE ? "oscope" : k_variation_names[v];
      const size_t n = strlen(vn);
      if (n > best_len && !strncmp(pname, vn, n) && pname[n] == '_') {
         best = v;
         best_len = n;
      }
   }
   return best;
}
/* flam3_copy_params (flam3.c:804-1030): the parameters of variation `var`, with the derived
 * values the reference copies along - and without flower_holes, which it forgets */
static void copy_var_params(flam3_xform *dst, const flam3_xform *src, int var) {
   for (size_t i = 0; i < sizeof(k_xform_params) / sizeof(k_xform_params[0]); i++) {
      if (param_variation(k_xform_params[i].name) != var) continue;
      if (!strcmp(k_xform_params[i].name, "flower_holes")) continue;
      *(double *)((char *)dst + k_xform_params[i].off) = *(const double *)((const char *)src + k_xform_params[i].off);
   }
   if (var == VAR_JULIAN) {
      dst->julian_rN = src->julian_rN;
      dst->julian_cn = src->julian_cn;
   } else if (var == VAR_JULIASCOPE) {
      dst->juliascope_rN = src->juliascope_rN;
      dst->juliascope_cn = src->juliascope_cn;
   } else if (var == VAR_PERSPECTIVE) {
      dst->persp_vsin = src->persp_vsin;
      dst->persp_vfcos = src->persp_vfcos;
   } else if (var == VAR_WEDGE_JULIA) {
      dst->wedgeJulia_cf = src->wedgeJulia_cf;
      dst->wedgeJulia_cn = src->wedgeJulia_cn;
      dst->wedgeJulia_rN = src->wedgeJulia_rN;
   }
}

/* flam3_copyx (flam3.c:1211-1268): a copy with nstd standard xforms (the missing ones are fresh
 * xforms marked as padding) and a final xform if any key frame has one */
static int copy_padded(flam3_genome *dst, const flam3_genome *src, int nstd, int want_final) {
   f3b_genome_init(dst, 1);
   memcpy(dst, src, sizeof(*dst));
   dst->xform = NULL;
   dst->chaos = NULL;
   dst->num_xforms = 0;
   dst->final_xform_index = -1;
   const int src_std = src->num_xforms - (src->final_xform_index >= 0);
   for (int i = 0; i < nstd; i++) {
      const int at = f3b_genome_add_xform(dst, 0);
      if (at < 0) return 1;
      dst->xform[at].padding = 1;
   }
   for (int i = 0; i < src_std; i++) {
      dst->xform[i] = src->xform[i]; /* (no motion elements in genomes of this front end) */
      dst->xform[i].motion = NULL;
      dst->xform[i].num_motion = 0;
      memcpy(dst->chaos[i], src->chaos[i], (size_t)src_std * sizeof(double));
   }
   if (want_final > 0) {
      const int at = f3b_genome_add_xform(dst, 1);
      if (at < 0) return 1;
      dst->xform[at].padding = 1;
      if (src->final_xform_enable > 0) {
         dst->xform[at] = src->xform[src->final_xform_index];
         dst->xform[at].motion = NULL;
         dst->xform[at].num_motion = 0;
      } else { /* a final xform to interpolate against: it must not rotate or recolour */
         dst->xform[at].animate = 0.0;
         dst->xform[at].color_speed = 0.0;
      }
   } else {
      dst->final_xform_index = -1;
      dst->final_xform_enable = 0;
   }
   return 0;
}

/* the better 'identity' of a padding xform (interpolation.c:845-1025): what it is made of
 * depends on what its neighbour in time uses */
static const int k_holes[] = {VAR_SPHERICAL, VAR_NGON, VAR_JULIAN, VAR_JULIASCOPE, VAR_POLAR, VAR_WEDGE_SPH, VAR_WEDGE_JULIA};

static void pad_like_neighbours(flam3_genome *g, int n, int i, int xf) {
   flam3_xform *x = &g[i].xform[xf];
   x->var[VAR_LINEAR] = 0.0;
   int found = 0;
   const int log_space = (i == 0 ? g[i].interpolation_type : g[i - 1].interpolation_type) == flam3_inttype_log;
   if (log_space)
      for (int d = -1; d <= 1; d += 2) {
         if (i + d < 0 || i + d >= n || g[i + d].xform[xf].padding == 1) continue;
         const flam3_xform *o = &g[i + d].xform[xf];
         int holes = 0;
         for (size_t k = 0; k < sizeof(k_holes) / sizeof(k_holes[0]); k++) holes |= o->var[k_holes[k]] > 0;
         if (holes) { /* against a half turn with weight -1 */
            x->var[VAR_LINEAR] = -1.0;
            x->c[0][0] = -1.0; x->c[0][1] = 0.0;
            x->c[1][0] = 0.0;  x->c[1][1] = -1.0;
            x->c[2][0] = 0.0;  x->c[2][1] = 0.0;
            found = -1;
         }
      }
   if (found == 0)
      for (int d = -1; d <= 1; d += 2) {
         if (i + d < 0 || i + d >= n || g[i + d].xform[xf].padding == 1) continue;
         const flam3_xform *o = &g[i + d].xform[xf];
         if (o->var[VAR_RECTANGLES] > 0) { x->var[VAR_RECTANGLES] = 1.0; x->rectangles_x = 0.0; x->rectangles_y = 0.0; found++; }
         if (o->var[VAR_RINGS2] > 0) { x->var[VAR_RINGS2] = 1.0; x->rings2_val = 0.0; found++; }
         if (o->var[VAR_FAN2] > 0) { x->var[VAR_FAN2] = 1.0; x->fan2_x = 0.0; x->fan2_y = 0.0; found++; }
         if (o->var[VAR_BLOB] > 0) { x->var[VAR_BLOB] = 1.0; x->blob_low = 1.0; x->blob_high = 1.0; x->blob_waves = 1.0; found++; }
         if (o->var[VAR_PERSPECTIVE] > 0) { x->var[VAR_PERSPECTIVE] = 1.0; x->perspective_angle = 0.0; found++; }
         if (o->var[VAR_CURL] > 0) { x->var[VAR_CURL] = 1.0; x->curl_c1 = 0.0; x->curl_c2 = 0.0; found++; }
         if (o->var[VAR_SUPER_SHAPE] > 0) {
            x->var[VAR_SUPER_SHAPE] = 1.0;
            x->super_shape_n1 = x->super_shape_n2 = x->super_shape_n3 = 2.0;
            x->super_shape_rnd = 0.0;
            x->super_shape_holes = 0.0;
            found++;
         }
      }
   if (found == 0) { /* fan and rings work on the coefficients: against the swapped axes */
      for (int d = -1; d <= 1; d += 2) {
         if (i + d < 0 || i + d >= n || g[i + d].xform[xf].padding == 1) continue;
         const flam3_xform *o = &g[i + d].xform[xf];
         if (o->var[VAR_FAN] > 0) { x->var[VAR_FAN] = 1.0; found++; }
         if (o->var[VAR_RINGS] > 0) { x->var[VAR_RINGS] = 1.0; found++; }
      }
      if (found > 0) {
         x->c[0][0] = 0.0; x->c[0][1] = 1.0;
         x->c[1][0] = 1.0; x->c[1][1] = 0.0;
         x->c[2][0] = 0.0; x->c[2][1] = 0.0;
      }
   }
   if (found == 0) x->var[VAR_LINEAR] = 1.0;
   else if (found > 0) {
      double sum = 0.0;
      for (int j = 0; j < flam3_nvariations; j++) sum += x->var[j];
      for (int j = 0; j < flam3_nvariations; j++) x->var[j] /= sum;
   }
}

/* flam3_align (interpolation.c:768-1032) */
static int align_genomes(flam3_genome *dst, const flam3_genome *src, int n) {
   int nstd = src[0].num_xforms - (src[0].final_xform_index >= 0), nfinal = src[0].final_xform_enable;
   int aligned = 1;
   for (int i = 1; i < n; i++) {
      const int t = src[i].num_xforms - (src[i].final_xform_index >= 0);
      if (t != nstd) {
         aligned = 0;
         if (t > nstd) nstd = t;
      }
      if (src[i].final_xform_enable != nfinal) {
         aligned = 0;
         nfinal |= src[i].final_xform_enable;
      }
   }
   for (int i = 0; i < n; i++)
      if (copy_padded(&dst[i], &src[i], nstd, nfinal)) return 1;
   const int lead = n / 2 - 1;
   if (dst[lead].interpolation_type == flam3_inttype_compat || dst[lead].interpolation_type == flam3_inttype_older) return 0;
   for (int i = 0; i < n; i++)
      for (int xf = 0; xf < nstd + nfinal; xf++) {
         /* a parametric variation that a neighbour in time uses and this xform does not: take
          * the neighbour's parameters, so that the blend does not sweep them from defaults */
         for (int j = 23; j < flam3_nvariations; j++) {
            if (dst[i].xform[xf].var[j] != 0) continue;
            if (i > 0) {
               if (dst[i - 1].xform[xf].var[j] != 0) copy_var_params(&dst[i].xform[xf], &dst[i - 1].xform[xf], j);
            } else if (i < n - 1) {
               if (dst[i + 1].xform[xf].var[j] != 0) copy_var_params(&dst[i].xform[xf], &dst[i + 1].xform[xf], j);
            }
         }
         if (dst[i].xform[xf].padding == 1 && !aligned) pad_like_neighbours(dst, n, i, xf);
      }
   return 0;
}

/* affine parts as two column vectors in polar form (interpolation.c:247-324) */
static void affine_to_polar(const flam3_genome *g, int n, int xf, int post, double ang[4][2], double mag[4][2], double trn[4][2]) {
   for (int k = 0; k < n; k++) {
      const double(*m)[2] = post ? g[k].xform[xf].post : g[k].xform[xf].c;
      int zero[2] = {0, 0};
      for (int col = 0; col < 2; col++) {
         ang[k][col] = atan2(m[col][1], m[col][0]);
         mag[k][col] = sqrt(m[col][0] * m[col][0] + m[col][1] * m[col][1]);
         if (mag[k][col] == 0.0) zero[col] = 1;
         trn[k][col] = m[2][col];
      }
      if (zero[0] == 1 && zero[1] == 0) ang[k][0] = ang[k][1];
      else if (zero[0] == 0 && zero[1] == 1) ang[k][1] = ang[k][0];
   }
   for (int col = 0; col < 2; col++)
      for (int k = 1; k < n; k++) {
         const double wind = g[k].xform[xf].wind[col];
         if (wind > 0 && !post) { /* asymmetric case: both angles within [wind - 2 pi, wind] */
            const double ref = wind - 2 * M_PI;
            while (ang[k - 1][col] < ref) ang[k - 1][col] += 2 * M_PI;
            while (ang[k - 1][col] > ref + 2 * M_PI) ang[k - 1][col] -= 2 * M_PI;
            while (ang[k][col] < ref) ang[k][col] += 2 * M_PI;
            while (ang[k][col] > ref + 2 * M_PI) ang[k][col] -= 2 * M_PI;
         } else { /* the shorter way round, clockwise at exactly half a turn */
            const double d = ang[k][col] - ang[k - 1][col];
            if (d > M_PI + 1e-10) ang[k][col] -= 2 * M_PI;
            else if (d < -(M_PI - 1e-10)) ang[k][col] += 2 * M_PI;
         }
      }
}
static void polar_blend(const double *c, int n, double ang[4][2], double mag[4][2], double trn[4][2], double out[3][2]) {
   double a[2] = {0.0, 0.0}, m[2] = {0.0, 0.0};
   int linear[2] = {0, 0};
   for (int col = 0; col < 2; col++)
      for (int i = 0; i < n; i++)
         if (log(mag[i][col]) < -10) linear[col] = 1; /* near-zero columns blend their lengths linearly */
   for (int i = 0; i < n; i++)
      for (int col = 0; col < 2; col++) {
         a[col] += c[i] * ang[i][col];
         if (!linear[col]) m[col] += c[i] * log(mag[i][col]);
         else m[col] += c[i] * mag[i][col];
         out[2][col] += c[i] * trn[i][col];
      }
   for (int col = 0; col < 2; col++) {
      const double len = linear[col] ? m[col] : exp(m[col]);
      out[col][0] = len * cos(a[col]);
      out[col][1] = len * sin(a[col]);
   }
}
static int is_identity(const double m[3][2]) {
   return m[0][0] == 1.0 && m[0][1] == 0.0 && m[1][0] == 0.0 && m[1][1] == 1.0 && m[2][0] == 0.0 && m[2][1] == 0.0;
}

/* flam3_interpolate_n (interpolation.c:373-708), stagger 0 */
#define BLEND(field)                                                         \
   do {                                                                      \
      double acc_ = 0.0;                                                     \
      for (int k_ = 0; k_ < n; k_++) acc_ += c[k_] * g[k_].field;            \
      r->field = acc_;                                                       \
   } while (0)
#define BLEND_INT(field)                                                     \
   do {                                                                      \
      double acc_ = 0.0;                                                     \
      for (int k_ = 0; k_ < n; k_++) acc_ += c[k_] * g[k_].field;            \
      r->field = (int)rint(acc_);                                            \
   } while (0)
/* get_stagger_coef (interpolation.c:343-367): with stagger the xforms do not blend together but
 * one after the other, the first xform first; t is the weight of the first key frame */
static double stagger_weight(double t, double stagger, int nxf, int xf) {
   const double max_stag = (double)(nxf - 1) / nxf;
   const double scaled = stagger * max_stag;
   const double st = scaled * (nxf - 1 - xf) / (nxf - 1);
   const double et = st + (1 - scaled);
   if (t <= st) return 0;
   if (t >= et) return 1;
   const double u = (t - st) / (1 - scaled);
   return (3.0 - (u + u)) * (u * u); /* 3u^2 - 2u^3 as the reference's -ffast-math build factors it */
}

static void blend_genomes(flam3_genome *r, int n, flam3_genome *g, const double *cw, double stagger) {
   double c[4];
   for (int k = 0; k < n; k++) c[k] = cw[k];
   if (g[0].palette_interpolation != flam3_palette_interpolation_sweep) {
      double rgb_fraction = 0.0;
      if (g[0].palette_interpolation == flam3_palette_interpolation_rgb) rgb_fraction = 1.0;
      else if (g[0].palette_interpolation == flam3_palette_interpolation_hsv_circular) rgb_fraction = g[0].hsv_rgb_palette_blend;
      for (int i = 0; i < 256; i++) {
         double new_rgb[3] = {0, 0, 0}, new_hsv[3] = {0, 0, 0}, count = 0, index = 0;
         int alpha1 = 1;
         for (int k = 0; k < n; k++) {
            double hsv[3];
            rgb_to_hsv(g[k].palette[i].color, hsv);
            if (n == 2 && k == 0 && g[0].palette_interpolation == flam3_palette_interpolation_hsv_circular) {
               double second[3];
               rgb_to_hsv(g[1].palette[i].color, second);
               if (second[0] - hsv[0] > 3.0) hsv[0] += 6.0;
               else if (second[0] - hsv[0] < -3.0) hsv[0] -= 6.0;
            }
            for (int j = 0; j < 3; j++) {
               new_rgb[j] += c[k] * g[k].palette[i].color[j];
               new_hsv[j] += c[k] * hsv[j];
            }
            count += c[k] * g[k].palette[i].color[3];
            if (g[k].palette[i].color[3] != 1.0) alpha1 = 0;
            index += c[k] * g[k].palette[i].index;
         }
         if (alpha1) count = 1.0;
         double back[3];
         hsv_to_rgb(new_hsv, back);
         double *out = r->palette[i].color;
         for (int j = 0; j < 3; j++) out[j] = rgb_fraction * new_rgb[j] + (1.0 - rgb_fraction) * back[j];
         out[3] = count;
         /* (the reference's -ffast-math build clips with max / min instructions: a -0.0 that
          * comes out of the blend leaves as +0.0) */
         for (int j = 0; j < 4; j++) {
            out[j] = out[j] > 0.0 ? out[j] : 0.0;
            out[j] = out[j] < 1.0 ? out[j] : 1.0;
         }
         index = index > 0.0 ? index : 0.0;
         r->palette[i].index = index < 255.0 ? index : 255.0;
      }
   } else
      for (int i = 0; i < 256; i++) r->palette[i] = g[i < 256 * c[0] ? 0 : 1].palette[i];

   r->palette_index = flam3_palette_random;
   r->symmetry = 0;
   r->spatial_filter_select = g[0].spatial_filter_select;
   r->temporal_filter_type = g[0].temporal_filter_type;
   r->palette_mode = g[0].palette_mode;
   r->interpolation_type = g[0].interpolation_type;
   r->palette_interpolation = g[0].palette_interpolation;
   r->hsv_rgb_palette_blend = g[0].hsv_rgb_palette_blend;
   BLEND(brightness); BLEND(contrast); BLEND(highlight_power); BLEND(gamma); BLEND(vibrancy); BLEND(hue_rotation);
   BLEND_INT(width); BLEND_INT(height); BLEND_INT(spatial_oversample);
   BLEND(center[0]); BLEND(center[1]); BLEND(rot_center[0]); BLEND(rot_center[1]);
   BLEND(background[0]); BLEND(background[1]); BLEND(background[2]);
   BLEND(pixels_per_unit); BLEND(spatial_filter_radius); BLEND(temporal_filter_exp); BLEND(temporal_filter_width);
   BLEND(sample_density); BLEND(zoom); BLEND(rotate);
   BLEND_INT(nbatches); BLEND_INT(ntemporal_samples);
   BLEND(estimator); BLEND(estimator_minimum); BLEND(estimator_curve); BLEND(gam_lin_thresh);

   const int nstd = g[0].num_xforms - (g[0].final_xform_index >= 0);
   for (int i = 0; i < nstd; i++)
      for (int j = 0; j < nstd; j++) {
         BLEND(chaos[i][j]);
         if (r->chaos[i][j] < 0) r->chaos[i][j] = 0;
      }
   for (int i = 0; i < g[0].num_xforms; i++) {
      if (n == 2 && stagger > 0 && i != g[0].final_xform_index) {
         c[0] = stagger_weight(cw[0], stagger, nstd, i);
         c[1] = 1.0 - c[0];
      } else {
         c[0] = cw[0];
         c[1] = cw[1];
      }
      BLEND(xform[i].density);
      if (r->xform[i].density < 0.0) r->xform[i].density = 0.0;
      BLEND(xform[i].color);
      if (r->xform[i].color < 0) r->xform[i].color = 0;
      if (r->xform[i].color > 1) r->xform[i].color = 1;
      BLEND(xform[i].color_speed);
      if (r->xform[i].color_speed < 0) r->xform[i].color_speed = 0;
      if (r->xform[i].color_speed > 1) r->xform[i].color_speed = 1;
      BLEND(xform[i].opacity);
      BLEND(xform[i].animate);
      for (size_t p = 0; p < sizeof(k_xform_params) / sizeof(k_xform_params[0]); p++) {
         double acc = 0.0;
         for (int k = 0; k < n; k++) acc += c[k] * *(const double *)((const char *)&g[k].xform[i] + k_xform_params[p].off);
         *(double *)((char *)&r->xform[i] + k_xform_params[p].off) = acc;
      }
      for (int j = 0; j < flam3_nvariations; j++) BLEND(xform[i].var[j]);

      int all_id = 1;
      for (int k = 0; k < n; k++) all_id &= is_identity(g[k].xform[i].post);
      memset(r->xform[i].c, 0, sizeof(r->xform[i].c));
      memset(r->xform[i].post, 0, sizeof(r->xform[i].post));
      if (g[0].interpolation_type == flam3_inttype_log) {
         double ang[4][2], mag[4][2], trn[4][2];
         affine_to_polar(g, n, i, 0, ang, mag, trn);
         polar_blend(c, n, ang, mag, trn, r->xform[i].c);
         if (!all_id) {
            affine_to_polar(g, n, i, 1, ang, mag, trn);
            polar_blend(c, n, ang, mag, trn, r->xform[i].post);
         }
      } else
         for (int k = 0; k < n; k++)
            for (int a = 0; a < 3; a++)
               for (int b = 0; b < 2; b++) {
                  r->xform[i].c[a][b] += c[k] * g[k].xform[i].c[a][b];
                  r->xform[i].post[a][b] += c[k] * g[k].xform[i].post[a][b];
               }
      if (all_id) {
         memset(r->xform[i].post, 0, sizeof(r->xform[i].post));
         r->xform[i].post[0][0] = 1.0;
         r->xform[i].post[1][1] = 1.0;
      }
   }
}

int f3b_interpolate(flam3_genome *cps, int ncps, double time, double stagger, flam3_genome *result) {
   if (ncps == 1) return f3b_genome_copy(result, &cps[0]);
   int i1, i2;
   if (cps[0].time >= time) { i1 = 0; i2 = 1; }
   else if (cps[ncps - 1].time <= time) { i1 = ncps - 2; i2 = ncps - 1; }
   else {
      i1 = 0;
      while (cps[i1].time < time) i1++;
      i1--;
      i2 = i1 + 1;
   }
   double c[4];
   c[0] = (cps[i2].time - time) / (cps[i2].time - cps[i1].time);
   c[1] = 1.0 - c[0];
   flam3_genome g[4];
   memset(g, 0, sizeof(g));
   int n = 2, smooth = 0;
   if (cps[i1].interpolation != flam3_interpolation_linear) {
      if (i1 == 0) fprintf(stderr, "error: cannot use smooth interpolation on first segment.\nreverting to linear interpolation.\n");
      else if (i2 == ncps - 1) fprintf(stderr, "error: cannot use smooth interpolation on last segment.\nreverting to linear interpolation.\n");
      else smooth = 1;
   }
   if (smooth) {
      n = 4;
      if (align_genomes(g, &cps[i1 - 1], 4)) return 1;
   } else if (align_genomes(g, &cps[i1], 2)) return 1;
   f3b_genome_init(result, 1);
   const int g_std = g[0].num_xforms - (g[0].final_xform_index >= 0);
   for (int i = 0; i < g_std; i++)
      if (f3b_genome_add_xform(result, 0) < 0) return 1;
   if (g[0].final_xform_index >= 0 && f3b_genome_add_xform(result, 1) < 0) return 1;
   result->time = time;
   result->interpolation = flam3_interpolation_linear;
   result->interpolation_type = g[0].interpolation_type;
   result->palette_interpolation = flam3_palette_interpolation_hsv_circular;
   result->hsv_rgb_palette_blend = 0.0;
   if (smooth) { /* Catmull-Rom weights of the four key frames (interpolation.c:326-337) */
      const double t = c[1], t2 = t * t, t3 = t2 * t;
      c[0] = (2 * t2 - t - t3) / 2;
      c[1] = (3 * t3 - 5 * t2 + 2) / 2;
      c[2] = ((4 * t2 + t) - 3 * t3) / 2; /* in this order: the reference's -ffast-math build reassociates the sum */
      c[3] = (t3 - t2) / 2;
   }
   blend_genomes(result, n, g, c, smooth ? 0.0 : stagger);
   for (int k = 0; k < n; k++) f3b_genome_free_xforms(&g[k]);
   return 0;
}


/* ---- writing genomes (flam3_print, flam3_print_xform, flam3.c:1674-2474) ---------------- */
/* The text flam3_print_to_string gives (no edit history, no 2.7 compatibility mode), character
 * for character for every fixture (tests/test_cpu_frontend.py). The parameters of a variation
 * are written when its weight is not zero, in the reference's order: */
static const struct {
   int var;
   const char *params; /* XML names, in print order */
} k_print_order[] = {
   {VAR_BLOB, "blob_low blob_high blob_waves"}, {VAR_PDJ, "pdj_a pdj_b pdj_c pdj_d"}, {VAR_FAN2, "fan2_x fan2_y"},
   {VAR_RINGS2, "rings2_val"}, {VAR_PERSPECTIVE, "perspective_angle perspective_dist"},
   {VAR_JULIAN, "julian_power julian_dist"}, {VAR_JULIASCOPE, "juliascope_power juliascope_dist"},
   {VAR_RADIAL_BLUR, "radial_blur_angle"}, {VAR_PIE, "pie_slices pie_rotation pie_thickness"},
   {VAR_NGON, "ngon_sides ngon_power ngon_corners ngon_circle"}, {VAR_CURL, "curl_c1 curl_c2"},
   {VAR_RECTANGLES, "rectangles_x rectangles_y"}, {VAR_DISC2, "disc2_rot disc2_twist"},
   {VAR_SUPER_SHAPE, "super_shape_rnd super_shape_m super_shape_n1 super_shape_n2 super_shape_n3 super_shape_holes"},
   {VAR_FLOWER, "flower_petals flower_holes"}, {VAR_CONIC, "conic_eccentricity conic_holes"},
   {VAR_PARABOLA, "parabola_height parabola_width"}, {VAR_BENT2, "bent2_x bent2_y"}, {VAR_BIPOLAR, "bipolar_shift"},
   {VAR_CELL, "cell_size"}, {VAR_CPOW, "cpow_i cpow_r cpow_power"},
   {VAR_CURVE, "curve_xamp curve_yamp curve_xlength curve_ylength"}, {VAR_ESCHER, "escher_beta"},
   {VAR_LAZYSUSAN, "lazysusan_x lazysusan_y lazysusan_spin lazysusan_space lazysusan_twist"},
   {VAR_MODULUS, "modulus_x modulus_y"},
   {VAR_OSCILLOSCOPE, "oscilloscope_separation oscilloscope_frequency oscilloscope_amplitude oscilloscope_damping"},
   {VAR_POPCORN2, "popcorn2_x popcorn2_y popcorn2_c"},
   /* (for an xform separation's four parameters are never written: flam3_print_xform forgets to
    * note that the variation is present, flam3.c:1945-2030 - a genome loses them on its way
    * through a file; motion elements do write them) */
   {-VAR_SEPARATION, "separation_x separation_y separation_xinside separation_yinside"},
   {VAR_SPLIT, "split_xsize split_ysize"}, {VAR_SPLITS, "splits_x splits_y"}, {VAR_STRIPES, "stripes_space stripes_warp"},
   {VAR_WEDGE, "wedge_angle wedge_hole wedge_count wedge_swirl"},
   {VAR_WEDGE_JULIA, "wedge_julia_angle wedge_julia_count wedge_julia_power wedge_julia_dist"},
   {VAR_WEDGE_SPH, "wedge_sph_angle wedge_sph_hole wedge_sph_count wedge_sph_swirl"},
   {VAR_WHORL, "whorl_inside whorl_outside"}, {VAR_WAVES2, "waves2_scalex waves2_scaley waves2_freqx waves2_freqy"},
   {VAR_AUGER, "auger_sym auger_weight auger_freq auger_scale"}, {VAR_FLUX, "flux_spread"},
   {VAR_MOBIUS, "mobius_re_a mobius_im_a mobius_re_b mobius_im_b mobius_re_c mobius_im_c mobius_re_d mobius_im_d"},
};

typedef struct {
   char *s;
   size_t n, cap;
   int failed;
} f3b_text;
static void text_add(f3b_text *t, const char *fmt, ...) {
   va_list ap;
   for (int pass = 0; pass < 2 && !t->failed; pass++) {
      va_start(ap, fmt);
      const int need = vsnprintf(t->s + t->n, t->cap - t->n, fmt, ap);
      va_end(ap);
      if (need < 0) {
         t->failed = 1;
         return;
      }
      if ((size_t)need < t->cap - t->n) {
         t->n += (size_t)need;
         return;
      }
      size_t cap = t->cap * 2 + (size_t)need + 64;
      char *grown = (char *)realloc(t->s, cap);
      if (!grown) {
         t->failed = 1;
         return;
      }
      t->s = grown;
      t->cap = cap;
   }
}
static const char *word_of(const f3b_word *words, int value) {
   for (; words->word; words++)
      if (words->value == value) return words->word;
   return NULL;
}
static double xform_param(const flam3_xform *x, const char *xml_name, size_t len) {
   char name[64];
   if (len >= sizeof(name)) return 0.0;
   memcpy(name, xml_name, len);
   name[len] = 0;
   const char *field = name;
   for (size_t a = 0; a < sizeof(k_xform_aliases) / sizeof(k_xform_aliases[0]); a++)
      if (!strcmp(name, k_xform_aliases[a].alias)) field = k_xform_aliases[a].name;
   for (size_t i = 0; i < sizeof(k_xform_params) / sizeof(k_xform_params[0]); i++)
      if (!strcmp(field, k_xform_params[i].name)) return *(const double *)((const char *)x + k_xform_params[i].off);
   return 0.0;
}
static int is_zero_matrix(const double m[3][2]) {
   return m[0][0] == 0.0 && m[0][1] == 0.0 && m[1][0] == 0.0 && m[1][1] == 0.0 && m[2][0] == 0.0 && m[2][1] == 0.0;
}
/* a <motion> element: only what is not zero, the parameters under their FIELD names with %f
 * (flam3.c:1896, 2291-2440) */
static void print_motion(f3b_text *t, const flam3_xform *x) {
   text_add(t, "      <motion motion_frequency=\"%d\" ", x->motion_freq);
   if (x->motion_func == MOTION_SIN) text_add(t, "motion_function=\"sin\" ");
   else if (x->motion_func == MOTION_TRIANGLE) text_add(t, "motion_function=\"triangle\" ");
   else if (x->motion_func == MOTION_HILL) text_add(t, "motion_function=\"hill\" ");
   if (x->color != 0.0) text_add(t, "color=\"%g\" ", x->color);
   if (x->color_speed != 0.0) text_add(t, "color_speed=\"%g\" ", x->color_speed);
   for (int j = 0; j < flam3_nvariations; j++)
      if (j != VAR_TWINTRIAN && x->var[j] != 0.0) text_add(t, "%s=\"%g\" ", k_variation_names[j], x->var[j]);
   for (size_t g = 0; g < sizeof(k_print_order) / sizeof(k_print_order[0]); g++)
      for (const char *p = k_print_order[g].params; *p;) {
         const size_t len = strcspn(p, " ");
         const double v = xform_param(x, p, len);
         if (v != 0.0) {
            if (!strncmp(p, "oscilloscope_", 13)) text_add(t, "oscope_%.*s=\"%f\" ", (int)(len - 13), p + 13, v);
            else text_add(t, "%.*s=\"%f\" ", (int)len, p, v);
         }
         p += len;
         while (*p == ' ') p++;
      }
   if (!is_zero_matrix(x->c))
      text_add(t, "coefs=\"%g %g %g %g %g %g\"", x->c[0][0], x->c[0][1], x->c[1][0], x->c[1][1], x->c[2][0], x->c[2][1]);
   if (!is_zero_matrix(x->post))
      text_add(t, " post=\"%g %g %g %g %g %g\"", x->post[0][0], x->post[0][1], x->post[1][0], x->post[1][1], x->post[2][0],
               x->post[2][1]);
   text_add(t, "/>\n");
}
static void print_xform(f3b_text *t, const flam3_xform *x, int is_final, int nstd, const double *chaos_row) {
   if (is_final) text_add(t, "   <finalxform ");
   else text_add(t, "   <xform weight=\"%g\" ", x->density);
   text_add(t, "color=\"%g\" color_speed=\"%g\" ", x->color, x->color_speed);
   if (!is_final) text_add(t, "animate=\"%g\" ", x->animate);
   for (int j = 0; j < flam3_nvariations; j++)
      if (j != VAR_TWINTRIAN && x->var[j] != 0.0) text_add(t, "%s=\"%g\" ", k_variation_names[j], x->var[j]);
   for (size_t g = 0; g < sizeof(k_print_order) / sizeof(k_print_order[0]); g++) {
      if (k_print_order[g].var < 0 || x->var[k_print_order[g].var] == 0.0) continue;
      for (const char *p = k_print_order[g].params; *p;) {
         const size_t len = strcspn(p, " ");
         text_add(t, "%.*s=\"%g\" ", (int)len, p, xform_param(x, p, len));
         p += len;
         while (*p == ' ') p++;
      }
   }
   text_add(t, "coefs=\"%g %g %g %g %g %g\"", x->c[0][0], x->c[0][1], x->c[1][0], x->c[1][1], x->c[2][0], x->c[2][1]);
   if (!is_identity(x->post))
      text_add(t, " post=\"%g %g %g %g %g %g\"", x->post[0][0], x->post[0][1], x->post[1][0], x->post[1][1], x->post[2][0],
               x->post[2][1]);
   if (!is_final) {
      int ncols = nstd;
      while (ncols > 0 && chaos_row[ncols - 1] == 1.0) ncols--;
      if (ncols > 0) {
         text_add(t, " chaos=\"");
         for (int j = 0; j < ncols; j++) text_add(t, "%g ", chaos_row[j]);
         text_add(t, "\"");
      }
   }
   text_add(t, " opacity=\"%g\"", x->opacity);
   if (x->num_motion > 0) {
      text_add(t, ">\n");
      for (int m = 0; m < x->num_motion; m++) print_motion(t, &x->motion[m]);
      text_add(t, is_final ? "   </finalxform>\n" : "   </xform>\n");
   } else
      text_add(t, "/>\n");
}

char *f3b_print_flam3(const flam3_genome *cp, const char *version, const char *extra_attributes) {
   f3b_text t = {(char *)malloc(1 << 14), 0, 1 << 14, 0};
   if (!t.s) return NULL;
   t.s[0] = 0;
   text_add(&t, "<flame version=\"%s\" time=\"%g\"", version ? version : "FLAM3-B200", cp->time);
   if (cp->flame_name[0]) text_add(&t, " name=\"%s\"", cp->flame_name);
   text_add(&t, " size=\"%d %d\" center=\"%g %g\" scale=\"%g\"", cp->width, cp->height, cp->center[0], cp->center[1],
            cp->pixels_per_unit);
   if (cp->zoom != 0.0) text_add(&t, " zoom=\"%g\"", cp->zoom);
   text_add(&t, " rotate=\"%g\" supersample=\"%d\" filter=\"%g\"", cp->rotate, cp->spatial_oversample, cp->spatial_filter_radius);
   const char *w = word_of(k_filter_shapes, cp->spatial_filter_select);
   if (w) text_add(&t, " filter_shape=\"%s\"", w);
   w = word_of(k_temporal, cp->temporal_filter_type);
   if (w) text_add(&t, " temporal_filter_type=\"%s\"", w);
   if (cp->temporal_filter_type == flam3_temporal_exp) text_add(&t, " temporal_filter_exp=\"%g\"", cp->temporal_filter_exp);
   text_add(&t, " temporal_filter_width=\"%g\" quality=\"%g\" passes=\"%d\" temporal_samples=\"%d\"", cp->temporal_filter_width,
            cp->sample_density, cp->nbatches, cp->ntemporal_samples);
   text_add(&t, " background=\"%g %g %g\" brightness=\"%g\" gamma=\"%g\" highlight_power=\"%g\" vibrancy=\"%g\"", cp->background[0],
            cp->background[1], cp->background[2], cp->brightness, cp->gamma, cp->highlight_power, cp->vibrancy);
   text_add(&t, " estimator_radius=\"%g\" estimator_minimum=\"%g\" estimator_curve=\"%g\" gamma_threshold=\"%g\"", cp->estimator,
            cp->estimator_minimum, cp->estimator_curve, cp->gam_lin_thresh);
   w = word_of(k_palette_mode, cp->palette_mode);
   if (w) text_add(&t, " palette_mode=\"%s\"", w);
   if (cp->interpolation != flam3_interpolation_linear) text_add(&t, " interpolation=\"smooth\"");
   w = cp->interpolation_type == flam3_inttype_compat ? "old" : word_of(k_interp_type, cp->interpolation_type);
   if (w) text_add(&t, " interpolation_type=\"%s\"", w);
   w = word_of(k_palette_interp, cp->palette_interpolation);
   if (w) text_add(&t, " palette_interpolation=\"%s\"", w);
   if (cp->palette_interpolation == flam3_palette_interpolation_hsv_circular && cp->hsv_rgb_palette_blend > 0.0)
      text_add(&t, " hsv_rgb_palette_blend=\"%g\"", cp->hsv_rgb_palette_blend);
   if (extra_attributes) text_add(&t, " %s", extra_attributes);
   text_add(&t, ">\n");
   if (cp->symmetry) text_add(&t, "   <symmetry kind=\"%d\"/>\n", cp->symmetry);
   const int nstd = cp->num_xforms - (cp->final_xform_index >= 0);
   for (int i = 0; i < cp->num_xforms; i++)
      print_xform(&t, &cp->xform[i], i == cp->final_xform_index, nstd, i == cp->final_xform_index ? NULL : cp->chaos[i]);
   for (int i = 0; i < 256; i++) {
      const double r = cp->palette[i].color[0] * 255.0, g = cp->palette[i].color[1] * 255.0;
      const double b = cp->palette[i].color[2] * 255.0, a = cp->palette[i].color[3] * 255.0;
      if (a == 255.0) text_add(&t, "   <color index=\"%d\" rgb=\"%.6g %.6g %.6g\"/>\n", i, r, g, b);
      else text_add(&t, "      <color index=\"%d\" rgba=\"%.6g %.6g %.6g %.6g\"/>\n", i, r, g, b, a);
   }
   text_add(&t, "</flame>\n");
   if (t.failed) {
      free(t.s);
      return NULL;
   }
   return t.s;
}

/* ---- output ----------------------------------------------------------------------------- */
int f3b_write_ppm(const char *path, const void *image, int width, int height, int nchan, int bytes_per_channel) {
   FILE *f = fopen(path, "wb");
   if (!f) {
      perror(path);
      return 1;
   }
   const int maxv = bytes_per_channel == 2 ? 65535 : 255;
   fprintf(f, "P6\n%d %d\n%d\n", width, height, maxv);
   const unsigned char *p = (const unsigned char *)image;
   for (size_t i = 0; i < (size_t)width * height; i++) {
      if (bytes_per_channel == 1) fwrite(p + i * nchan, 1, 3, f);
      else { /* PPM wants the most significant byte first */
         const unsigned short *q = (const unsigned short *)(p + i * nchan * 2);
         for (int c = 0; c < 3; c++) {
            const unsigned char be[2] = {(unsigned char)(q[c] >> 8), (unsigned char)(q[c] & 255)};
            fwrite(be, 1, 2, f);
         }
      }
   }
   return fclose(f) != 0;
}
