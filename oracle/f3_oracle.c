/*
 * oracle/f3_oracle.c — TEST INFRASTRUCTURE: a plain-C, double-precision, single-threaded
 * restatement of flam3's render hot path, used only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker of the CUDA path. The product never links it.
 *
 * PARITY PINNING: every function here is checked in tests/test_cpu_oracle.py against the
 * UNMODIFIED reference compiled from /root/reference (oracle/_ref/libflam3_ref.so) — the
 * reference ships no golden vectors of its own — and against the fixtures
 * those runs produced (tests/golden/ npz files, generator tools/make_golden.py):
 *   ISAAC streams bit for bit, chooser tables entry for entry, apply_xform for all 99
 *   variations to 1e-9 relative, filter banks to 1e-12, density estimation / tone map to
 *   float rounding, final 8-bit images bit for bit on the reference's own accumulator.
 *
 * What it restates (reference file:line):
 *   isaac / irandinit / irand                 isaac.c:25-110, isaac.h:49-52 (64-bit ub4 quirk)
 *   flam3_random_isaac_01/11/bit              flam3.c:2519-2544
 *   flam3_create_xform_distrib / chaos_distrib flam3.c:76-224
 *   prepare_precalc_flags / xform_precalc     variations.c:1942-2126
 *   apply_xform + 99 variations               variations.c:142-1937, 2129-2399
 *   flam3_iterate                             flam3.c:232-299
 *   iter_thread's bucket bump                 rect.c:411-516
 *   spatial / temporal / DE filter creation   filters.c:47-489
 *   de_thread (scatter order)                 rect.c:57-178
 *   no-DE log scale, early clip, final filter rect.c:947-974, 1086-1247
 *   calc_alpha / calc_newrgb / rgb<->hsv      palettes.c:199-349
 *
 * The variation maths is written once for a packed parameter vector (o_pack_params) so the
 * host-side derived constants of the reference's *_precalc functions are explicit.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define O_EPS 1e-10
#define O_NVARS 99
#define O_GRAIN 16384

/* ------------------------------------------------------------------------------------ */
/* ISAAC as the reference runs it on LP64                                                 */
/* ------------------------------------------------------------------------------------ */
typedef unsigned long o_ub4; /* isaacs.h:17 — 64-bit containers, results masked to 32 */
typedef struct {
   o_ub4 randcnt, randrsl[16], randmem[16], randa, randb, randc;
} o_rand;
#define O_M32(x) ((x)&0xfffffffful)

void o_isaac(o_rand *c) {
   o_ub4 a = c->randa, b = O_M32(c->randb + (++c->randc));
   for (int i = 0; i < 16; i++) {
      const o_ub4 x = c->randmem[i];
      o_ub4 s;
      switch (i & 3) {
      case 0: s = a << 13; break;
      case 1: s = a >> 6; break;
      case 2: s = a << 2; break;
      default: s = a >> 16; break;
      }
      a = O_M32((a ^ s) + c->randmem[(i + 8) & 15]);
      const o_ub4 y = O_M32(c->randmem[(x >> 2) & 15] + a + b);
      c->randmem[i] = y;
      b = O_M32(c->randmem[((y >> 4) >> 2) & 15] + x);
      c->randrsl[i] = b;
   }
   c->randb = b;
   c->randa = a;
}

static void o_mix8(o_ub4 *v) { /* isaac.c:48-59, NOT masked: the LP64 quirk */
   v[0] ^= v[1] << 11; v[3] += v[0]; v[1] += v[2];
   v[1] ^= v[2] >> 2;  v[4] += v[1]; v[2] += v[3];
   v[2] ^= v[3] << 8;  v[5] += v[2]; v[3] += v[4];
   v[3] ^= v[4] >> 16; v[6] += v[3]; v[4] += v[5];
   v[4] ^= v[5] << 10; v[7] += v[4]; v[5] += v[6];
   v[5] ^= v[6] >> 4;  v[0] += v[5]; v[6] += v[7];
   v[6] ^= v[7] << 8;  v[1] += v[6]; v[7] += v[0];
   v[7] ^= v[0] >> 9;  v[2] += v[7]; v[0] += v[1];
}

void o_irandinit(o_rand *c, int flag) {
   o_ub4 v[8];
   c->randa = c->randb = c->randc = 0;
   for (int i = 0; i < 8; i++) v[i] = 0x9e3779b9ul;
   for (int i = 0; i < 4; i++) o_mix8(v);
   if (flag) {
      for (int pass = 0; pass < 2; pass++)
         for (int i = 0; i < 16; i += 8) {
            const o_ub4 *src = pass ? c->randmem : c->randrsl;
            for (int k = 0; k < 8; k++) v[k] += src[i + k];
            o_mix8(v);
            for (int k = 0; k < 8; k++) c->randmem[i + k] = v[k];
         }
   } else {
      for (int i = 0; i < 16; i += 8) {
         o_mix8(v);
         for (int k = 0; k < 8; k++) c->randmem[i + k] = v[k];
      }
   }
   o_isaac(c);
   c->randcnt = 16;
}

o_ub4 o_irand(o_rand *c) {
   if (c->randcnt-- == 0) {
      o_isaac(c);
      c->randcnt = 15;
   }
   return c->randrsl[c->randcnt];
}

void o_seed(o_rand *c, const char *seed) { /* flam3_init_frame with isaac_seed set */
   memset(c, 0, sizeof(*c));
   strncpy((char *)c->randrsl, seed, 16 * sizeof(o_ub4));
   o_irandinit(c, 1);
}

static double o_u01(o_rand *c) { return ((int)o_irand(c) & 0xfffffff) / (double)0xfffffff; }
static double o_u11(o_rand *c) {
   return (((int)o_irand(c) & 0xfffffff) - 0x7ffffff) / (double)0x7ffffff;
}
static int o_bit(o_rand *c) { return (int)o_irand(c) & 1; }
static double o_gauss4(o_rand *c) {
   const double a = o_u01(c), b = o_u01(c), d = o_u01(c), e = o_u01(c);
   return a + b + d + e - 2.0;
}
static int o_bad(double x) { return (x != x) || (x > 1e10) || (x < -1e10); }

/* ------------------------------------------------------------------------------------ */
/* genome model                                                                           */
/* ------------------------------------------------------------------------------------ */
#define O_PARAMS(X)                                                                          \
   X(blob_low) X(blob_high) X(blob_waves) X(pdj_a) X(pdj_b) X(pdj_c) X(pdj_d) X(fan2_x)       \
   X(fan2_y) X(rings2_val) X(perspective_angle) X(perspective_dist) X(julian_power)          \
   X(julian_dist) X(juliascope_power) X(juliascope_dist) X(radial_blur_angle) X(pie_slices)  \
   X(pie_rotation) X(pie_thickness) X(ngon_sides) X(ngon_power) X(ngon_circle)               \
   X(ngon_corners) X(curl_c1) X(curl_c2) X(rectangles_x) X(rectangles_y) X(disc2_rot)        \
   X(disc2_twist) X(super_shape_rnd) X(super_shape_m) X(super_shape_n1) X(super_shape_n2)    \
   X(super_shape_n3) X(super_shape_holes) X(flower_petals) X(flower_holes)                   \
   X(conic_eccentricity) X(conic_holes) X(parabola_height) X(parabola_width) X(bent2_x)      \
   X(bent2_y) X(bipolar_shift) X(cell_size) X(cpow_r) X(cpow_i) X(cpow_power) X(curve_xamp)  \
   X(curve_yamp) X(curve_xlength) X(curve_ylength) X(escher_beta) X(lazysusan_spin)          \
   X(lazysusan_space) X(lazysusan_twist) X(lazysusan_x) X(lazysusan_y) X(modulus_x)          \
   X(modulus_y) X(oscope_separation) X(oscope_frequency) X(oscope_amplitude)                 \
   X(oscope_damping) X(popcorn2_x) X(popcorn2_y) X(popcorn2_c) X(separation_x)               \
   X(separation_xinside) X(separation_y) X(separation_yinside) X(split_xsize) X(split_ysize) \
   X(splits_x) X(splits_y) X(stripes_space) X(stripes_warp) X(wedge_angle) X(wedge_hole)     \
   X(wedge_count) X(wedge_swirl) X(wedge_julia_angle) X(wedge_julia_count)                   \
   X(wedge_julia_power) X(wedge_julia_dist) X(wedge_sph_angle) X(wedge_sph_count)            \
   X(wedge_sph_hole) X(wedge_sph_swirl) X(whorl_inside) X(whorl_outside) X(waves2_freqx)     \
   X(waves2_scalex) X(waves2_freqy) X(waves2_scaley) X(auger_sym) X(auger_weight)            \
   X(auger_freq) X(auger_scale) X(flux_spread) X(mobius_re_a) X(mobius_im_a) X(mobius_re_b)  \
   X(mobius_im_b) X(mobius_re_c) X(mobius_im_c) X(mobius_re_d) X(mobius_im_d)

typedef struct {
   double var[O_NVARS];
   double c[3][2], post[3][2];
   double density, color, color_speed, opacity;
#define X(n) double n;
   O_PARAMS(X)
#undef X
} o_xform;

static const char *const o_var_names[O_NVARS] = {
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

enum {
   OV_LINEAR, OV_SINUSOIDAL, OV_SPHERICAL, OV_SWIRL, OV_HORSESHOE, OV_POLAR, OV_HANDKERCHIEF,
   OV_HEART, OV_DISC, OV_SPIRAL, OV_HYPERBOLIC, OV_DIAMOND, OV_EX, OV_JULIA, OV_BENT, OV_WAVES,
   OV_FISHEYE, OV_POPCORN, OV_EXPONENTIAL, OV_POWER, OV_COSINE, OV_RINGS, OV_FAN, OV_BLOB, OV_PDJ,
   OV_FAN2, OV_RINGS2, OV_EYEFISH, OV_BUBBLE, OV_CYLINDER, OV_PERSPECTIVE, OV_NOISE, OV_JULIAN,
   OV_JULIASCOPE, OV_BLUR, OV_GAUSSIAN_BLUR, OV_RADIAL_BLUR, OV_PIE, OV_NGON, OV_CURL,
   OV_RECTANGLES, OV_ARCH, OV_TANGENT, OV_SQUARE, OV_RAYS, OV_BLADE, OV_SECANT2, OV_TWINTRIAN,
   OV_CROSS, OV_DISC2, OV_SUPER_SHAPE, OV_FLOWER, OV_CONIC, OV_PARABOLA, OV_BENT2, OV_BIPOLAR,
   OV_BOARDERS, OV_BUTTERFLY, OV_CELL, OV_CPOW, OV_CURVE, OV_EDISC, OV_ELLIPTIC, OV_ESCHER, OV_FOCI,
   OV_LAZYSUSAN, OV_LOONIE, OV_PRE_BLUR, OV_MODULUS, OV_OSCILLOSCOPE, OV_POLAR2, OV_POPCORN2,
   OV_SCRY, OV_SEPARATION, OV_SPLIT, OV_SPLITS, OV_STRIPES, OV_WEDGE, OV_WEDGE_JULIA, OV_WEDGE_SPH,
   OV_WHORL, OV_WAVES2, OV_EXP, OV_LOG, OV_SIN, OV_COS, OV_TAN, OV_SEC, OV_CSC, OV_COT, OV_SINH,
   OV_COSH, OV_TANH, OV_SECH, OV_CSCH, OV_COTH, OV_AUGER, OV_FLUX, OV_MOBIUS
};

typedef struct {
   int num_xforms;        /* including the final xform if present */
   int final_xform_index; /* -1 if none; must be the last xform */
   int final_xform_enable;
   o_xform *xform;
   double *chaos; /* num_std x num_std, row = previous xform; NULL = all ones */
} o_genome;

o_genome *o_genome_new(int num_xforms, int final_index, int final_enable) {
   o_genome *g = (o_genome *)calloc(1, sizeof(o_genome));
   g->num_xforms = num_xforms;
   g->final_xform_index = final_index;
   g->final_xform_enable = final_enable;
   g->xform = (o_xform *)calloc((size_t)num_xforms, sizeof(o_xform));
   for (int i = 0; i < num_xforms; i++) {
      o_xform *x = &g->xform[i];
      x->c[0][0] = x->c[1][1] = x->post[0][0] = x->post[1][1] = 1.0;
      x->opacity = 1.0;
      x->color_speed = 0.5;
   }
   return g;
}
void o_genome_free(o_genome *g) {
   if (!g) return;
   free(g->xform);
   free(g->chaos);
   free(g);
}
void o_genome_set_chaos(o_genome *g, const double *m, int n) {
   free(g->chaos);
   g->chaos = (double *)malloc(sizeof(double) * (size_t)n * n);
   memcpy(g->chaos, m, sizeof(double) * (size_t)n * n);
}
/* Set one named quantity of xform xi: a variation name (its weight), a parameter name as in
 * the .flam3 format, "weight", "color", "color_speed", "opacity", or coefs/post by index
 * "c00".."c21" / "p00".."p21". Returns 0 if the name is known. */
int o_xform_set(o_genome *g, int xi, const char *name, double v) {
   o_xform *x = &g->xform[xi];
   for (int i = 0; i < O_NVARS; i++)
      if (!strcmp(name, o_var_names[i])) { x->var[i] = v; return 0; }
#define X(n) if (!strcmp(name, #n)) { x->n = v; return 0; }
   O_PARAMS(X)
#undef X
   if (!strcmp(name, "weight")) { x->density = v; return 0; }
   if (!strcmp(name, "color")) { x->color = v; return 0; }
   if (!strcmp(name, "color_speed")) { x->color_speed = v; return 0; }
   if (!strcmp(name, "opacity")) { x->opacity = v; return 0; }
   if ((name[0] == 'c' || name[0] == 'p') && strlen(name) == 3) {
      const int i = name[1] - '0', j = name[2] - '0';
      if (i >= 0 && i < 3 && j >= 0 && j < 2) {
         if (name[0] == 'c') x->c[i][j] = v; else x->post[i][j] = v;
         return 0;
      }
   }
   return 1;
}
const char *o_var_name(int i) { return (i >= 0 && i < O_NVARS) ? o_var_names[i] : NULL; }

/* packed parameters of a variation, with the reference's *_precalc constants folded in */
static int o_pack_params(const o_xform *x, int id, double *p) {
   switch (id) {
   case OV_WAVES: { /* waves_precalc, variations.c:1969 */
      const double dx = x->c[2][0], dy = x->c[2][1];
      p[0] = x->c[1][0];
      p[1] = 1.0 / (dx * dx + O_EPS);
      p[2] = x->c[1][1];
      p[3] = 1.0 / (dy * dy + O_EPS);
      return 4;
   }
   case OV_POPCORN: p[0] = x->c[2][0]; p[1] = x->c[2][1]; return 2;
   case OV_RINGS: p[0] = x->c[2][0] * x->c[2][0] + O_EPS; return 1;
   case OV_FAN:
      p[0] = M_PI * (x->c[2][0] * x->c[2][0] + O_EPS);
      p[1] = x->c[2][1];
      p[2] = 0.5 * p[0];
      return 3;
   case OV_BLOB:
      p[0] = x->blob_low; p[1] = x->blob_high - x->blob_low; p[2] = x->blob_waves;
      return 3;
   case OV_PDJ: p[0] = x->pdj_a; p[1] = x->pdj_b; p[2] = x->pdj_c; p[3] = x->pdj_d; return 4;
   case OV_FAN2:
      p[0] = M_PI * (x->fan2_x * x->fan2_x + O_EPS);
      p[1] = x->fan2_y;
      p[2] = 0.5 * p[0];
      return 3;
   case OV_RINGS2: p[0] = x->rings2_val * x->rings2_val + O_EPS; return 1;
   case OV_PERSPECTIVE: { /* perspective_precalc, variations.c:1942 */
      const double ang = x->perspective_angle * M_PI / 2.0;
      p[0] = x->perspective_dist;
      p[1] = sin(ang);
      p[2] = x->perspective_dist * cos(ang);
      return 3;
   }
   case OV_JULIAN: /* juliaN_precalc, variations.c:1948 */
      p[0] = x->julian_power;
      p[1] = x->julian_dist / x->julian_power / 2.0;
      p[2] = fabs(x->julian_power);
      return 3;
   case OV_JULIASCOPE: /* juliaScope_precalc, variations.c:1959 */
      p[0] = x->juliascope_power;
      p[1] = x->juliascope_dist / x->juliascope_power / 2.0;
      p[2] = fabs(x->juliascope_power);
      return 3;
   case OV_RADIAL_BLUR: /* radial_blur_precalc, variations.c:1964 */
      p[0] = sin(x->radial_blur_angle * M_PI / 2.0);
      p[1] = cos(x->radial_blur_angle * M_PI / 2.0);
      return 2;
   case OV_PIE: p[0] = x->pie_slices; p[1] = x->pie_rotation; p[2] = x->pie_thickness; return 3;
   case OV_NGON:
      p[0] = x->ngon_power / 2.0;
      p[1] = 2.0 * M_PI / x->ngon_sides;
      p[2] = x->ngon_corners;
      p[3] = x->ngon_circle;
      return 4;
   case OV_CURL: p[0] = x->curl_c1; p[1] = x->curl_c2; return 2;
   case OV_RECTANGLES: p[0] = x->rectangles_x; p[1] = x->rectangles_y; return 2;
   case OV_DISC2: { /* disc2_precalc, variations.c:1977 */
      const double add = x->disc2_twist;
      double s = sin(add), c = cos(add) - 1.0;
      if (add > 2 * M_PI) {
         const double k = 1 + add - 2 * M_PI;
         c *= k;
         s *= k;
      }
      if (add < -2 * M_PI) {
         const double k = 1 + add + 2 * M_PI;
         c *= k;
         s *= k;
      }
      p[0] = x->disc2_rot * M_PI;
      p[1] = s;
      p[2] = c;
      return 3;
   }
   case OV_SUPER_SHAPE: /* supershape_precalc, variations.c:1999 */
      p[0] = x->super_shape_m / 4.0;
      p[1] = x->super_shape_n2;
      p[2] = x->super_shape_n3;
      p[3] = x->super_shape_rnd;
      p[4] = x->super_shape_holes;
      p[5] = -1.0 / x->super_shape_n1;
      return 6;
   case OV_FLOWER: p[0] = x->flower_petals; p[1] = x->flower_holes; return 2;
   case OV_CONIC: p[0] = x->conic_eccentricity; p[1] = x->conic_holes; return 2;
   case OV_PARABOLA: p[0] = x->parabola_height; p[1] = x->parabola_width; return 2;
   case OV_BENT2: p[0] = x->bent2_x; p[1] = x->bent2_y; return 2;
   case OV_BIPOLAR: p[0] = -(M_PI / 2.0) * x->bipolar_shift; return 1;
   case OV_CELL: p[0] = x->cell_size; p[1] = 1.0 / x->cell_size; return 2;
   case OV_CPOW:
      p[0] = 2.0 * M_PI / x->cpow_power;
      p[1] = x->cpow_r / x->cpow_power;
      p[2] = x->cpow_i / x->cpow_power;
      p[3] = x->cpow_power;
      return 4;
   case OV_CURVE: {
      double xl = x->curve_xlength * x->curve_xlength, yl = x->curve_ylength * x->curve_ylength;
      if (xl < 1e-20) xl = 1e-20;
      if (yl < 1e-20) yl = 1e-20;
      p[0] = x->curve_xamp;
      p[1] = x->curve_yamp;
      p[2] = 1.0 / xl;
      p[3] = 1.0 / yl;
      return 4;
   }
   case OV_ESCHER:
      p[0] = 0.5 * (1.0 + cos(x->escher_beta));
      p[1] = 0.5 * sin(x->escher_beta);
      return 2;
   case OV_LAZYSUSAN:
      p[0] = x->lazysusan_spin; p[1] = x->lazysusan_space; p[2] = x->lazysusan_twist;
      p[3] = x->lazysusan_x; p[4] = x->lazysusan_y;
      return 5;
   case OV_MODULUS: p[0] = x->modulus_x; p[1] = x->modulus_y; return 2;
   case OV_OSCILLOSCOPE:
      p[0] = 2.0 * M_PI * x->oscope_frequency;
      p[1] = x->oscope_amplitude;
      p[2] = x->oscope_separation;
      p[3] = x->oscope_damping;
      return 4;
   case OV_POPCORN2: p[0] = x->popcorn2_x; p[1] = x->popcorn2_y; p[2] = x->popcorn2_c; return 3;
   case OV_SEPARATION:
      p[0] = x->separation_x * x->separation_x;
      p[1] = x->separation_xinside;
      p[2] = x->separation_y * x->separation_y;
      p[3] = x->separation_yinside;
      return 4;
   case OV_SPLIT: p[0] = x->split_xsize * M_PI; p[1] = x->split_ysize * M_PI; return 2;
   case OV_SPLITS: p[0] = x->splits_x; p[1] = x->splits_y; return 2;
   case OV_STRIPES: p[0] = 1.0 - x->stripes_space; p[1] = x->stripes_warp; return 2;
   case OV_WEDGE:
      p[0] = x->wedge_angle; p[1] = x->wedge_hole; p[2] = x->wedge_count; p[3] = x->wedge_swirl;
      p[4] = 1.0 - x->wedge_angle * x->wedge_count * M_1_PI * 0.5;
      return 5;
   case OV_WEDGE_JULIA: /* wedgeJulia_precalc, variations.c:1953 */
      p[0] = x->wedge_julia_dist / x->wedge_julia_power / 2.0;
      p[1] = fabs(x->wedge_julia_power);
      p[2] = x->wedge_julia_power;
      p[3] = x->wedge_julia_count;
      p[4] = 1.0 - x->wedge_julia_angle * x->wedge_julia_count * M_1_PI * 0.5;
      p[5] = x->wedge_julia_angle;
      return 6;
   case OV_WEDGE_SPH:
      p[0] = x->wedge_sph_angle; p[1] = x->wedge_sph_hole; p[2] = x->wedge_sph_count;
      p[3] = x->wedge_sph_swirl;
      p[4] = 1.0 - x->wedge_sph_angle * x->wedge_sph_count * M_1_PI * 0.5;
      return 5;
   case OV_WHORL: p[0] = x->whorl_inside; p[1] = x->whorl_outside; return 2;
   case OV_WAVES2:
      p[0] = x->waves2_freqx; p[1] = x->waves2_scalex; p[2] = x->waves2_freqy;
      p[3] = x->waves2_scaley;
      return 4;
   case OV_AUGER:
      p[0] = x->auger_sym; p[1] = x->auger_weight; p[2] = x->auger_freq; p[3] = x->auger_scale;
      return 4;
   case OV_FLUX: p[0] = 2.0 + x->flux_spread; return 1;
   case OV_MOBIUS:
      p[0] = x->mobius_re_a; p[1] = x->mobius_im_a; p[2] = x->mobius_re_b; p[3] = x->mobius_im_b;
      p[4] = x->mobius_re_c; p[5] = x->mobius_im_c; p[6] = x->mobius_re_d; p[7] = x->mobius_im_d;
      return 8;
   default:
      return 0;
   }
}


/* interpolation.c:22-29 */
static double o_vis_adjust(double opacity) {
   return opacity == 0.0 ? 0.0 : pow(10.0, -log(1.0 / opacity) / log(2.0));
}

static int o_in(int id, const int *set, int n) {
   for (int i = 0; i < n; i++)
      if (set[i] == id) return 1;
   return 0;
}

/*
 * apply_xform, variations.c:2129-2399. p = {x, y, color, vis} in, q out; returns 1 on a bad
 * value (q[0..1] replaced by fresh uniform draws).
 */
int o_apply_xform(const o_xform *x, const double *pin, double *q, o_rand *rc) {
   static const int need_atan_xy[] = {5, 6, 7, 8, 12, 13, 22, 23, 25, 49, 70};
   static const int need_angles[] = {9, 10, 11, 19, 21, 23, 26};
   static const int need_atan_yx[] = {32, 33, 36, 38, 50, 51, 52, 59, 63, 77, 78, 79, 80, 83};
   const double s1 = x->color_speed;
   q[2] = s1 * x->color + (1.0 - s1) * pin[2];
   q[3] = o_vis_adjust(x->opacity);

   double tx = x->c[0][0] * pin[0] + x->c[1][0] * pin[1] + x->c[2][0];
   double ty = x->c[0][1] * pin[0] + x->c[1][1] * pin[1] + x->c[2][1];

   if (x->var[OV_PRE_BLUR] != 0.0) { /* var67_pre_blur */
      const double g = x->var[OV_PRE_BLUR] * o_gauss4(rc);
      const double a = o_u01(rc) * 2.0 * M_PI;
      tx += g * cos(a);
      ty += g * sin(a);
   }

   const double sumsq = tx * tx + ty * ty;
   const double rad = sqrt(sumsq);
   double atan_xy = 0.0, atan_yx = 0.0, sina = 0.0, cosa = 0.0;
   int fa = 0, fb = 0, fc = 0;
   for (int id = 0; id < O_NVARS; id++)
      if (x->var[id] != 0.0) {
         fa |= o_in(id, need_atan_xy, 11);
         fb |= o_in(id, need_angles, 7);
         fc |= o_in(id, need_atan_yx, 14);
      }
   if (fa) atan_xy = atan2(tx, ty);
   if (fb) {
      sina = tx / rad;
      cosa = ty / rad;
   }
   if (fc) atan_yx = atan2(ty, tx);

   double p0 = 0.0, p1 = 0.0;
   for (int id = 0; id < O_NVARS; id++) {
      const double w = x->var[id];
      if (w == 0.0 || id == OV_PRE_BLUR) continue;
      double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      o_pack_params(x, id, p);
      switch (id) {
      case 0: /* linear */
         p0 += w * tx;
         p1 += w * ty;
         break;
      case 2: { /* spherical */
         const double r2 = w / (sumsq + O_EPS);
         p0 += r2 * tx;
         p1 += r2 * ty;
      } break;
   case 1: /* sinusoidal */
      p0 += w * sin(tx);
      p1 += w * sin(ty);
      break;
   case 3: { /* swirl */
      double s, c;
      sincos(sumsq, &s, &c);
      p0 += w * (s * tx - c * ty);
      p1 += w * (c * tx + s * ty);
   } break;
   case 4: { /* horseshoe */
      double r = w / (rad + O_EPS);
      p0 += (tx - ty) * (tx + ty) * r;
      p1 += 2.0 * tx * ty * r;
   } break;
   case 5: /* polar */
      p0 += w * (atan_xy * M_1_PI);
      p1 += w * (rad - 1.0);
      break;
   case 6: /* handkerchief */
      p0 += w * rad * sin(atan_xy + rad);
      p1 += w * rad * cos(atan_xy - rad);
      break;
   case 7: { /* heart */
      double s, c;
      sincos(rad * atan_xy, &s, &c);
      double r = w * rad;
      p0 += r * s;
      p1 -= r * c;
   } break;
   case 8: { /* disc */
      double a = atan_xy * M_1_PI, s, c;
      sincos(M_PI * rad, &s, &c);
      p0 += w * s * a;
      p1 += w * c * a;
   } break;
   case 9: { /* spiral */
      double r = rad + O_EPS, s, c;
      double r1 = w / r;
      sincos(r, &s, &c);
      p0 += r1 * (cosa + s);
      p1 += r1 * (sina - c);
   } break;
   case 10: { /* hyperbolic */
      double r = rad + O_EPS;
      p0 += w * sina / r;
      p1 += w * cosa * r;
   } break;
   case 11: { /* diamond */
      double s, c;
      sincos(rad, &s, &c);
      p0 += w * sina * c;
      p1 += w * cosa * s;
   } break;
   case 12: { /* ex */
      double n0 = sin(atan_xy + rad);
      double n1 = cos(atan_xy - rad);
      double m0 = n0 * n0 * n0 * rad;
      double m1 = n1 * n1 * n1 * rad;
      p0 += w * (m0 + m1);
      p1 += w * (m0 - m1);
   } break;
   case 13: { /* julia */
      double a = 0.5 * atan_xy, s, c;
      if (o_bit(rc)) a += M_PI;
      double r = w * sqrt(rad);
      sincos(a, &s, &c);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 14: { /* bent */
      double nx = tx < 0.0 ? tx * 2.0 : tx;
      double ny = ty < 0.0 ? ty * 0.5 : ty;
      p0 += w * nx;
      p1 += w * ny;
   } break;
   case 15: { /* waves: p = {c10, dx2, c11, dy2} */
      p0 += w * (tx + p[0] * sin(ty * p[1]));
      p1 += w * (ty + p[2] * sin(tx * p[3]));
   } break;
   case 16: { /* fisheye (x/y swap is intentional in the reference) */
      double r = 2.0 * w / (rad + 1.0);
      p0 += r * ty;
      p1 += r * tx;
   } break;
   case 17: { /* popcorn: p = {c20, c21} */
      double dx = tan(3.0 * ty);
      double dy = tan(3.0 * tx);
      p0 += w * (tx + p[0] * sin(dx));
      p1 += w * (ty + p[1] * sin(dy));
   } break;
   case 18: { /* exponential */
      double dx = w * exp(tx - 1.0), s, c;
      sincos(M_PI * ty, &s, &c);
      p0 += dx * c;
      p1 += dx * s;
   } break;
   case 19: { /* power */
      double r = w * pow(rad, sina);
      p0 += r * cosa;
      p1 += r * sina;
   } break;
   case 20: { /* cosine */
      double s, c;
      sincos(tx * M_PI, &s, &c);
      p0 += w * (c * cosh(ty));
      p1 -= w * (s * sinh(ty));
   } break;
   case 21: { /* rings: p = {dx} */
      double dx = p[0];
      double r = w * (fmod(rad + dx, 2.0 * dx) - dx + rad * (1.0 - dx));
      p0 += r * cosa;
      p1 += r * sina;
   } break;
   case 22: { /* fan: p = {dx, dy, dx2} */
      double a = atan_xy, s, c;
      a += (fmod(a + p[1], p[0]) > p[2]) ? -p[2] : p[2];
      sincos(a, &s, &c);
      double r = w * rad;
      p0 += r * c;
      p1 += r * s;
   } break;
   case 23: { /* blob: p = {low, high-low, waves} */
      double r = rad * (p[0] + p[1] * (0.5 + 0.5 * sin(p[2] * atan_xy)));
      p0 += w * sina * r;
      p1 += w * cosa * r;
   } break;
   case 24: { /* pdj: p = {a,b,c,d} */
      double nx1 = cos(p[1] * tx);
      double nx2 = sin(p[2] * tx);
      double ny1 = sin(p[0] * ty);
      double ny2 = cos(p[3] * ty);
      p0 += w * (ny1 - nx1);
      p1 += w * (nx2 - ny2);
   } break;
   case 25: { /* fan2: p = {dx, dy, dx2} */
      double a = atan_xy, s, c;
      double t = a + p[1] - p[0] * trunc((a + p[1]) / p[0]);
      a = (t > p[2]) ? a - p[2] : a + p[2];
      sincos(a, &s, &c);
      double r = w * rad;
      p0 += r * s;
      p1 += r * c;
   } break;
   case 26: { /* rings2: p = {dx} */
      double dx = p[0];
      double r = rad;
      r += -2.0 * dx * trunc((r + dx) / (2.0 * dx)) + r * (1.0 - dx);
      p0 += w * sina * r;
      p1 += w * cosa * r;
   } break;
   case 27: { /* eyefish */
      double r = (w * 2.0) / (rad + 1.0);
      p0 += r * tx;
      p1 += r * ty;
   } break;
   case 28: { /* bubble */
      double r = w / (0.25 * sumsq + 1.0);
      p0 += r * tx;
      p1 += r * ty;
   } break;
   case 29: /* cylinder */
      p0 += w * sin(tx);
      p1 += w * ty;
      break;
   case 30: { /* perspective: p = {dist, vsin, vfcos} */
      double t = 1.0 / (p[0] - ty * p[1]);
      p0 += w * p[0] * tx * t;
      p1 += w * p[2] * ty * t;
   } break;
   case 31: { /* noise */
      double s, c;
      sincos(o_u01(rc) * 2.0 * M_PI, &s, &c);
      double r = w * o_u01(rc);
      p0 += tx * r * c;
      p1 += ty * r * s;
   } break;
   case 32: { /* julian: p = {power, cn, rN} */
      double t_rnd = trunc(p[2] * o_u01(rc));
      double a = (atan_yx + 2.0 * M_PI * t_rnd) / p[0], s, c;
      double r = w * pow(sumsq, p[1]);
      sincos(a, &s, &c);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 33: { /* juliascope: p = {power, cn, rN} */
      int t_rnd = (int)trunc(p[2] * o_u01(rc));
      double a, s, c;
      if ((t_rnd & 1) == 0) a = (2.0 * M_PI * (double)t_rnd + atan_yx) / p[0];
      else a = (2.0 * M_PI * (double)t_rnd - atan_yx) / p[0];
      sincos(a, &s, &c);
      double r = w * pow(sumsq, p[1]);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 34: { /* blur */
      double s, c;
      sincos(o_u01(rc) * 2.0 * M_PI, &s, &c);
      double r = w * o_u01(rc);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 35: { /* gaussian_blur */
      double s, c;
      sincos(o_u01(rc) * 2.0 * M_PI, &s, &c);
      double r = w * o_gauss4(rc);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 36: { /* radial_blur: p = {spinvar, zoomvar} */
      double g = w * o_gauss4(rc), s, c;
      sincos(atan_yx + p[0] * g, &s, &c);
      double rz = p[1] * g - 1.0;
      p0 += rad * c + rz * tx;
      p1 += rad * s + rz * ty;
   } break;
   case 37: { /* pie: p = {slices, rotation, thickness} */
      int sl = (int)(o_u01(rc) * p[0] + 0.5);
      double a = p[1] + 2.0 * M_PI * ((double)sl + o_u01(rc) * p[2]) / p[0], s, c;
      double r = w * o_u01(rc);
      sincos(a, &s, &c);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 38: { /* ngon: p = {power/2, b=2pi/sides, corners, circle} */
      double r_factor = pow(sumsq, p[0]);
      double b = p[1];
      double phi = atan_yx - b * floor(atan_yx / b);
      if (phi > 0.5 * b) phi -= b;
      double amp = p[2] * (1.0 / (cos(phi) + O_EPS) - 1.0) + p[3];
      amp /= (r_factor + O_EPS);
      p0 += w * tx * amp;
      p1 += w * ty * amp;
   } break;
   case 39: { /* curl: p = {c1, c2} */
      double re = 1.0 + p[0] * tx + p[1] * (tx * tx - ty * ty);
      double im = p[0] * ty + 2.0 * p[1] * tx * ty;
      double r = w / (re * re + im * im);
      p0 += (tx * re + ty * im) * r;
      p1 += (ty * re - tx * im) * r;
   } break;
   case 40: { /* rectangles: p = {x, y} */
      if (p[0] == 0.0) p0 += w * tx;
      else p0 += w * ((2.0 * floor(tx / p[0]) + 1.0) * p[0] - tx);
      if (p[1] == 0.0) p1 += w * ty;
      else p1 += w * ((2.0 * floor(ty / p[1]) + 1.0) * p[1] - ty);
   } break;
   case 41: { /* arch (weight used non-linearly, as in the reference) */
      double s, c;
      sincos(o_u01(rc) * w * M_PI, &s, &c);
      p0 += w * s;
      p1 += w * (s * s) / c;
   } break;
   case 42: /* tangent */
      p0 += w * sin(tx) / cos(ty);
      p1 += w * tan(ty);
      break;
   case 43: { /* square */
      double a = o_u01(rc);
      double b = o_u01(rc);
      p0 += w * (a - 0.5);
      p1 += w * (b - 0.5);
   } break;
   case 44: { /* rays */
      double ang = w * o_u01(rc) * M_PI;
      double r = w / (sumsq + O_EPS);
      double tanr = w * tan(ang) * r;
      p0 += tanr * cos(tx);
      p1 += tanr * sin(ty);
   } break;
   case 45: { /* blade */
      double r = o_u01(rc) * w * rad, s, c;
      sincos(r, &s, &c);
      p0 += w * tx * (c + s);
      p1 += w * tx * (c - s);
   } break;
   case 46: { /* secant2 */
      double cr = cos(w * rad);
      double icr = 1.0 / cr;
      p0 += w * tx;
      p1 += (cr < 0.0) ? w * (icr + 1.0) : w * (icr - 1.0);
   } break;
   case 47: { /* twintrian */
      double r = o_u01(rc) * w * rad, s, c;
      sincos(r, &s, &c);
      double diff = log10(s * s) + c;
      if (o_bad(diff)) diff = -30.0;
      p0 += w * tx * diff;
      p1 += w * tx * (diff - s * M_PI);
   } break;
   case 48: { /* cross */
      double s = tx * tx - ty * ty;
      double r = w * sqrt(1.0 / (s * s + O_EPS));
      p0 += tx * r;
      p1 += ty * r;
   } break;
   case 49: { /* disc2: p = {timespi, sinadd, cosadd} */
      double s, c;
      sincos(p[0] * (tx + ty), &s, &c);
      double r = w * atan_xy * M_1_PI;
      p0 += (s + p[2]) * r;
      p1 += (c + p[1]) * r;
   } break;
   case 50: { /* super_shape: p = {pm_4, n2, n3, rnd, holes, pneg1_n1} */
      double st, ct;
      sincos(p[0] * atan_yx + M_PI_4, &st, &ct);
      double t1 = pow(fabs(ct), p[1]);
      double t2 = pow(fabs(st), p[2]);
      double r = w * ((p[3] * o_u01(rc) + (1.0 - p[3]) * rad) - p[4]) * pow(t1 + t2, p[5]) /
                rad;
      p0 += r * tx;
      p1 += r * ty;
   } break;
   case 51: { /* flower: p = {petals, holes} */
      double r = w * (o_u01(rc) - p[1]) * cos(p[0] * atan_yx) / rad;
      p0 += r * tx;
      p1 += r * ty;
   } break;
   case 52: { /* conic: p = {eccentricity, holes} */
      double ct = tx / rad;
      double r = w * (o_u01(rc) - p[1]) * p[0] / (1.0 + p[0] * ct) / rad;
      p0 += r * tx;
      p1 += r * ty;
   } break;
   case 53: { /* parabola: p = {height, width} */
      double s, c;
      sincos(rad, &s, &c);
      double a = o_u01(rc);
      double b = o_u01(rc);
      p0 += p[0] * w * s * s * a;
      p1 += p[1] * w * c * b;
   } break;
   case 54: { /* bent2: p = {x, y} */
      double nx = tx < 0.0 ? tx * p[0] : tx;
      double ny = ty < 0.0 ? ty * p[1] : ty;
      p0 += w * nx;
      p1 += w * ny;
   } break;
   case 55: { /* bipolar: p = {ps = -pi/2 * shift} */
      double t = sumsq + 1.0;
      double x2 = 2.0 * tx;
      double y = 0.5 * atan2(2.0 * ty, sumsq - 1.0) + p[0];
      if (y > M_PI_2) y = -M_PI_2 + fmod(y + M_PI_2, M_PI);
      else if (y < -M_PI_2) y = M_PI_2 - fmod(M_PI_2 - y, M_PI);
      p0 += w * 0.25 * M_2_PI * log((t + x2) / (t - x2));
      p1 += w * M_2_PI * y;
   } break;
   case 56: { /* boarders */
      double rx = rint(tx), ry = rint(ty);
      double ox = tx - rx, oy = ty - ry;
      if (o_u01(rc) >= 0.75) {
         p0 += w * (ox * 0.5 + rx);
         p1 += w * (oy * 0.5 + ry);
      } else if (fabs(ox) >= fabs(oy)) {
         if (ox >= 0.0) {
            p0 += w * (ox * 0.5 + rx + 0.25);
            p1 += w * (oy * 0.5 + ry + 0.25 * oy / ox);
         } else {
            p0 += w * (ox * 0.5 + rx - 0.25);
            p1 += w * (oy * 0.5 + ry - 0.25 * oy / ox);
         }
      } else {
         if (oy >= 0.0) {
            p1 += w * (oy * 0.5 + ry + 0.25);
            p0 += w * (ox * 0.5 + rx + ox / oy * 0.25);
         } else {
            p1 += w * (oy * 0.5 + ry - 0.25);
            p0 += w * (ox * 0.5 + rx - ox / oy * 0.25);
         }
      }
   } break;
   case 57: { /* butterfly; 4/sqrt(3*pi) */
      double wx = w * 1.3029400317411197908970256609023;
      double y2 = ty * 2.0;
      double r = wx * sqrt(fabs(ty * tx) / (O_EPS + tx * tx + y2 * y2));
      p0 += r * tx;
      p1 += r * y2;
   } break;
   case 58: { /* cell: p = {size, 1/size} */
      int x = (int)floor(tx * p[1]);
      int y = (int)floor(ty * p[1]);
      double dx = tx - (double)x * p[0];
      double dy = ty - (double)y * p[0];
      if (y >= 0) {
         if (x >= 0) { y *= 2; x *= 2; }
         else { y *= 2; x = -(2 * x + 1); }
      } else {
         if (x >= 0) { y = -(2 * y + 1); x *= 2; }
         else { y = -(2 * y + 1); x = -(2 * x + 1); }
      }
      p0 += w * (dx + (double)x * p[0]);
      p1 -= w * (dy + (double)y * p[0]);
   } break;
   case 59: { /* cpow: p = {va, vc, vd, power} */
      double lnr = 0.5 * log(sumsq), s, c;
      double ang = p[1] * atan_yx + p[2] * lnr + p[0] * floor(p[3] * o_u01(rc));
      double m = w * exp(p[1] * lnr - p[2] * atan_yx);
      sincos(ang, &s, &c);
      p0 += m * c;
      p1 += m * s;
   } break;
   case 60: { /* curve: p = {xamp, yamp, 1/xlen^2, 1/ylen^2} */
      p0 += w * (tx + p[0] * exp(-ty * ty * p[2]));
      p1 += w * (ty + p[1] * exp(-tx * tx * p[3]));
   } break;
   case 61: { /* edisc */
      double tmp = sumsq + 1.0;
      double tmp2 = 2.0 * tx;
      double xmax = (sqrt(tmp + tmp2) + sqrt(tmp - tmp2)) * 0.5;
      double a1 = log(xmax + sqrt(xmax - 1.0));
      double a2 = -acos(tx / xmax);
      double ww = w / 11.57034632, snv, csv;
      sincos(a1, &snv, &csv);
      double snhu = sinh(a2), cshu = cosh(a2);
      if (ty > 0.0) snv = -snv;
      p0 += ww * cshu * csv;
      p1 += ww * snhu * snv;
   } break;
   case 62: { /* elliptic */
      double tmp = sumsq + 1.0;
      double x2 = 2.0 * tx;
      double xmax = 0.5 * (sqrt(tmp + x2) + sqrt(tmp - x2));
      double a = tx / xmax;
      double b = 1.0 - a * a;
      double ssx = xmax - 1.0;
      double ww = w / M_PI_2;
      b = b < 0.0 ? 0.0 : sqrt(b);
      ssx = ssx < 0.0 ? 0.0 : sqrt(ssx);
      p0 += ww * atan2(a, b);
      if (ty > 0.0) p1 += ww * log(xmax + ssx);
      else p1 -= ww * log(xmax + ssx);
   } break;
   case 63: { /* escher: p = {vc, vd} */
      double lnr = 0.5 * log(sumsq), s, c;
      double m = w * exp(p[0] * lnr - p[1] * atan_yx);
      double n = p[0] * atan_yx + p[1] * lnr;
      sincos(n, &s, &c);
      p0 += m * c;
      p1 += m * s;
   } break;
   case 64: { /* foci */
      double expx = exp(tx) * 0.5;
      double expnx = 0.25 / expx, s, c;
      sincos(ty, &s, &c);
      double t = w / (expx + expnx - c);
      p0 += t * (expx - expnx);
      p1 += t * s;
   } break;
   case 65: { /* lazysusan: p = {spin, space, twist, x, y} */
      double x = tx - p[3];
      double y = ty + p[4];
      double r = sqrt(x * x + y * y);
      if (r < w) {
         double a = atan2(y, x) + p[0] + p[2] * (w - r), s, c;
         sincos(a, &s, &c);
         r = w * r;
         p0 += r * c + p[3];
         p1 += r * s - p[4];
      } else {
         r = w * (1.0 + p[1] / r);
         p0 += r * x + p[3];
         p1 += r * y - p[4];
      }
   } break;
   case 66: { /* loonie */
      double w2 = w * w;
      if (sumsq < w2) {
         double r = w * sqrt(w2 / sumsq - 1.0);
         p0 += r * tx;
         p1 += r * ty;
      } else {
         p0 += w * tx;
         p1 += w * ty;
      }
   } break;
   /* 67 pre_blur is applied before the loop and never packed as a loop entry */
   case 68: { /* modulus: p = {x, y} */
      double xr = 2.0 * p[0], yr = 2.0 * p[1];
      if (tx > p[0]) p0 += w * (-p[0] + fmod(tx + p[0], xr));
      else if (tx < -p[0]) p0 += w * (p[0] - fmod(p[0] - tx, xr));
      else p0 += w * tx;
      if (ty > p[1]) p1 += w * (-p[1] + fmod(ty + p[1], yr));
      else if (ty < -p[1]) p1 += w * (p[1] - fmod(p[1] - ty, yr));
      else p1 += w * ty;
   } break;
   case 69: { /* oscilloscope: p = {2*pi*freq, amplitude, separation, damping} */
      double t;
      if (p[3] == 0.0) t = p[1] * cos(p[0] * tx) + p[2];
      else t = p[1] * exp(-fabs(tx) * p[3]) * cos(p[0] * tx) + p[2];
      p0 += w * tx;
      if (fabs(ty) <= t) p1 -= w * ty;
      else p1 += w * ty;
   } break;
   case 70: { /* polar2 */
      double p2v = w * M_1_PI;
      p0 += p2v * atan_xy;
      p1 += p2v * 0.5 * log(sumsq);
   } break;
   case 71: { /* popcorn2: p = {x, y, c} */
      p0 += w * (tx + p[0] * sin(tan(ty * p[2])));
      p1 += w * (ty + p[1] * sin(tan(tx * p[2])));
   } break;
   case 72: { /* scry */
      double r = 1.0 / (rad * (sumsq + 1.0 / (w + O_EPS)));
      p0 += tx * r;
      p1 += ty * r;
   } break;
   case 73: { /* separation: p = {x^2, xinside, y^2, yinside} */
      if (tx > 0.0) p0 += w * (sqrt(tx * tx + p[0]) - tx * p[1]);
      else p0 -= w * (sqrt(tx * tx + p[0]) + tx * p[1]);
      if (ty > 0.0) p1 += w * (sqrt(ty * ty + p[2]) - ty * p[3]);
      else p1 -= w * (sqrt(ty * ty + p[2]) + ty * p[3]);
   } break;
   case 74: { /* split: p = {xsize*pi, ysize*pi} */
      if (cos(tx * p[0]) >= 0.0) p1 += w * ty;
      else p1 -= w * ty;
      if (cos(ty * p[1]) >= 0.0) p0 += w * tx;
      else p0 -= w * tx;
   } break;
   case 75: { /* splits: p = {x, y} */
      p0 += w * (tx >= 0.0 ? tx + p[0] : tx - p[0]);
      p1 += w * (ty >= 0.0 ? ty + p[1] : ty - p[1]);
   } break;
   case 76: { /* stripes: p = {1-space, warp} */
      double roundx = floor(tx + 0.5);
      double offx = tx - roundx;
      p0 += w * (offx * p[0] + roundx);
      p1 += w * (ty + offx * offx * p[1]);
   } break;
   case 77: { /* wedge: p = {angle, hole, count, swirl, comp_fac} */
      double a = atan_yx + p[3] * rad, s, c;
      double cc = floor((p[2] * a + M_PI) * M_1_PI * 0.5);
      a = a * p[4] + cc * p[0];
      sincos(a, &s, &c);
      double r = w * (rad + p[1]);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 78: { /* wedge_julia: p = {cn, rN, power, count, cf, angle} */
      double r = w * pow(sumsq, p[0]), s, c;
      int t_rnd = (int)(p[1] * o_u01(rc));
      double a = (atan_yx + 2.0 * M_PI * (double)t_rnd) / p[2];
      double cc = floor((p[3] * a + M_PI) * M_1_PI * 0.5);
      a = a * p[4] + cc * p[5];
      sincos(a, &s, &c);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 79: { /* wedge_sph: p = {angle, hole, count, swirl, comp_fac} */
      double r = 1.0 / (rad + O_EPS), s, c;
      double a = atan_yx + p[3] * r;
      double cc = floor((p[2] * a + M_PI) * M_1_PI * 0.5);
      a = a * p[4] + cc * p[0];
      sincos(a, &s, &c);
      r = w * (r + p[1]);
      p0 += r * c;
      p1 += r * s;
   } break;
   case 80: { /* whorl: p = {inside, outside} */
      double a, s, c;
      if (rad < w) a = atan_yx + p[0] / (w - rad);
      else a = atan_yx + p[1] / (w - rad);
      sincos(a, &s, &c);
      p0 += w * rad * c;
      p1 += w * rad * s;
   } break;
   case 81: { /* waves2: p = {freqx, scalex, freqy, scaley} */
      p0 += w * (tx + p[1] * sin(ty * p[0]));
      p1 += w * (ty + p[3] * sin(tx * p[2]));
   } break;
   case 82: { /* exp */
      double e = exp(tx), s, c;
      sincos(ty, &s, &c);
      p0 += w * e * c;
      p1 += w * e * s;
   } break;
   case 83: /* log */
      p0 += w * 0.5 * log(sumsq);
      p1 += w * atan_yx;
      break;
   case 84: { /* sin */
      double s, c;
      sincos(tx, &s, &c);
      p0 += w * s * cosh(ty);
      p1 += w * c * sinh(ty);
   } break;
   case 85: { /* cos */
      double s, c;
      sincos(tx, &s, &c);
      p0 += w * c * cosh(ty);
      p1 -= w * s * sinh(ty);
   } break;
   case 86: { /* tan */
      double s, c;
      sincos(2.0 * tx, &s, &c);
      double den = 1.0 / (c + cosh(2.0 * ty));
      p0 += w * den * s;
      p1 += w * den * sinh(2.0 * ty);
   } break;
   case 87: { /* sec */
      double s, c;
      sincos(tx, &s, &c);
      double den = 2.0 / (cos(2.0 * tx) + cosh(2.0 * ty));
      p0 += w * den * c * cosh(ty);
      p1 += w * den * s * sinh(ty);
   } break;
   case 88: { /* csc */
      double s, c;
      sincos(tx, &s, &c);
      double den = 2.0 / (cosh(2.0 * ty) - cos(2.0 * tx));
      p0 += w * den * s * cosh(ty);
      p1 -= w * den * c * sinh(ty);
   } break;
   case 89: { /* cot */
      double s, c;
      sincos(2.0 * tx, &s, &c);
      double den = 1.0 / (cosh(2.0 * ty) - c);
      p0 += w * den * s;
      p1 -= w * den * sinh(2.0 * ty);
   } break;
   case 90: { /* sinh */
      double s, c;
      sincos(ty, &s, &c);
      p0 += w * sinh(tx) * c;
      p1 += w * cosh(tx) * s;
   } break;
   case 91: { /* cosh */
      double s, c;
      sincos(ty, &s, &c);
      p0 += w * cosh(tx) * c;
      p1 += w * sinh(tx) * s;
   } break;
   case 92: { /* tanh */
      double s, c;
      sincos(2.0 * ty, &s, &c);
      double den = 1.0 / (c + cosh(2.0 * tx));
      p0 += w * den * sinh(2.0 * tx);
      p1 += w * den * s;
   } break;
   case 93: { /* sech */
      double s, c;
      sincos(ty, &s, &c);
      double den = 2.0 / (cos(2.0 * ty) + cosh(2.0 * tx));
      p0 += w * den * c * cosh(tx);
      p1 -= w * den * s * sinh(tx);
   } break;
   case 94: { /* csch */
      double s, c;
      sincos(ty, &s, &c);
      double den = 2.0 / (cosh(2.0 * tx) - cos(2.0 * ty));
      p0 += w * den * sinh(tx) * c;
      p1 -= w * den * cosh(tx) * s;
   } break;
   case 95: { /* coth */
      double s, c;
      sincos(2.0 * ty, &s, &c);
      double den = 1.0 / (cosh(2.0 * tx) - c);
      p0 += w * den * sinh(2.0 * tx);
      p1 += w * den * s;
   } break;
   case 96: { /* auger: p = {sym, weight, freq, scale} */
      double s = sin(p[2] * tx);
      double t = sin(p[2] * ty);
      double dy = ty + p[1] * (p[3] * s * 0.5 + fabs(ty) * s);
      double dx = tx + p[1] * (p[3] * t * 0.5 + fabs(tx) * t);
      p0 += w * (tx + p[0] * (dx - tx));
      p1 += w * dy;
   } break;
   case 97: { /* flux: p = {2 + spread} */
      double xpw = tx + w, xmw = tx - w;
      double avgr = w * p[0] * sqrt(sqrt(ty * ty + xpw * xpw) / sqrt(ty * ty + xmw * xmw));
      double avga = (atan2(ty, xmw) - atan2(ty, xpw)) * 0.5, s, c;
      sincos(avga, &s, &c);
      p0 += avgr * c;
      p1 += avgr * s;
   } break;
   case 98: { /* mobius: p = {re_a, im_a, re_b, im_b, re_c, im_c, re_d, im_d} */
      double re_u = p[0] * tx - p[1] * ty + p[2];
      double im_u = p[0] * ty + p[1] * tx + p[3];
      double re_v = p[4] * tx - p[5] * ty + p[6];
      double im_v = p[4] * ty + p[5] * tx + p[7];
      double rv = w / (re_v * re_v + im_v * im_v);
      p0 += rv * (re_u * re_v + im_u * im_v);
      p1 += rv * (im_u * re_v - re_u * im_v);
   } break;
      default:
         break;
      }
   }

   const int has_post = !(x->post[0][0] == 1.0 && x->post[0][1] == 0.0 && x->post[1][0] == 0.0 &&
                          x->post[1][1] == 1.0 && x->post[2][0] == 0.0 && x->post[2][1] == 0.0);
   if (has_post) {
      q[0] = x->post[0][0] * p0 + x->post[1][0] * p1 + x->post[2][0];
      q[1] = x->post[0][1] * p0 + x->post[1][1] * p1 + x->post[2][1];
   } else {
      q[0] = p0;
      q[1] = p1;
   }
   if (o_bad(q[0]) || o_bad(q[1])) {
      q[0] = o_u11(rc);
      q[1] = o_u11(rc);
      return 1;
   }
   return 0;
}

/* ------------------------------------------------------------------------------------ */
/* xform chooser (flam3.c:76-224)                                                         */
/* ------------------------------------------------------------------------------------ */
static int o_num_std(const o_genome *g) { return g->num_xforms - (g->final_xform_index >= 0); }

int o_chaos_enabled(const o_genome *g) {
   const int n = o_num_std(g);
   if (!g->chaos) return 0;
   for (int i = 0; i < n * n; i++)
      if (fabs(g->chaos[i] - 1.0) > O_EPS) return 1;
   return 0;
}

static int o_chooser_row(const o_genome *g, int xi, unsigned short *row) {
   const int n = o_num_std(g);
   double dr = 0.0;
   for (int i = 0; i < n; i++) {
      double d = g->xform[i].density;
      if (xi >= 0) d *= g->chaos[xi * n + i];
      if (d < 0.0) return 1;
      dr += d;
   }
   if (dr == 0.0) return 1;
   dr /= O_GRAIN;
   int j = 0;
   double t = g->xform[0].density * (xi >= 0 ? g->chaos[xi * n] : 1.0);
   double r = 0.0;
   for (int i = 0; i < O_GRAIN; i++) {
      while (r >= t) {
         j++;
         t += g->xform[j].density * (xi >= 0 ? g->chaos[xi * n + j] : 1.0);
      }
      row[i] = (unsigned short)j;
      r += dr;
   }
   return 0;
}

/* out must hold (num_std + 1) * 16384 entries; returns the number of rows, 0 on error */
int o_chooser_table(const o_genome *g, unsigned short *out) {
   const int n = o_num_std(g);
   if (o_chooser_row(g, -1, out)) return 0;
   if (!o_chaos_enabled(g)) return 1;
   for (int i = 0; i < n; i++)
      if (o_chooser_row(g, i, out + (size_t)(i + 1) * O_GRAIN)) return 0;
   return n + 1;
}

/* ------------------------------------------------------------------------------------ */
/* flam3_iterate (flam3.c:232-299): samples[0..3] is the start point                      */
/* ------------------------------------------------------------------------------------ */
int o_iterate(const o_genome *g, int n, int fuse, double *samples, const unsigned short *table,
              o_rand *rc) {
   const int xaos = o_chaos_enabled(g);
   double p[4] = {samples[0], samples[1], samples[2], samples[3]}, q[4];
   int consec = 0, bad = 0, lastxf = 0;
   for (int i = -4 * fuse; i < 4 * n; i += 4) {
      const unsigned k = (unsigned)o_irand(rc) & (O_GRAIN - 1);
      const int fn = xaos ? table[(size_t)lastxf * O_GRAIN + k] : table[k];
      if (o_apply_xform(&g->xform[fn], p, q, rc) > 0) {
         consec++;
         bad++;
         if (consec < 5) {
            memcpy(p, q, sizeof(p));
            i -= 4;
            continue;
         }
         consec = 0;
      } else
         consec = 0;
      lastxf = fn + 1;
      memcpy(p, q, sizeof(p));
      if (g->final_xform_enable == 1) {
         const o_xform *fx = &g->xform[g->final_xform_index];
         if (fx->opacity == 1 || o_u01(rc) < fx->opacity) {
            o_apply_xform(fx, p, q, rc);
            q[3] = p[3];
         }
      }
      if (i >= 0) memcpy(samples + i, q, sizeof(q));
   }
   return bad;
}

/* ------------------------------------------------------------------------------------ */
/* bucket bump (rect.c:411-516), bits=33 semantics: unsigned buckets, truncating adds     */
/* ------------------------------------------------------------------------------------ */
typedef struct {
   double bounds[4], rot[2][2], rot_center[2], ws0, wb0s0, hs1, hb1s1;
   int rotate, width, height, cmap_size, palette_linear;
} o_camera;

static void o_bump(uint32_t *dest, double delta) { /* flam3.c:3838-3840 */
   if ((double)(0xffffffffu - *dest) > delta) *dest = (uint32_t)((double)*dest + delta);
   else *dest = 0xffffffffu;
}

/* dmap: 256 x {r,g,b,a} already scaled by 255 * temporal weight (rect.c:778-782) */
void o_accumulate(const double *samples, int n, const o_camera *cam, const double *dmap,
                  uint32_t *buckets) {
   for (int j = 0; j < n; j++) {
      const double *p = samples + 4 * j;
      double p0 = p[0], p1 = p[1];
      if (cam->rotate) {
         const double a = p[0] - cam->rot_center[0], b = p[1] - cam->rot_center[1];
         p0 = a * cam->rot[0][0] + b * cam->rot[0][1] + cam->rot_center[0];
         p1 = a * cam->rot[1][0] + b * cam->rot[1][1] + cam->rot_center[1];
      }
      if (!(p0 >= cam->bounds[0] && p1 >= cam->bounds[1] && p0 <= cam->bounds[2] &&
            p1 <= cam->bounds[3]))
         continue;
      if (p[3] == 0) continue;
      const int ix = (int)(cam->ws0 * p0 - cam->wb0s0), iy = (int)(cam->hs1 * p1 - cam->hb1s1);
      if (ix < 0 || iy < 0 || ix >= cam->width || iy >= cam->height) continue; /* guards the
                        reference's inclusive upper bound ("bugs not to copy") */
      uint32_t *b = buckets + 5 * ((size_t)ix + (size_t)cam->width * iy);
      const double di = p[2] * cam->cmap_size;
      int ci = (int)di;
      double col[4];
      if (cam->palette_linear) {
         double frac;
         if (ci < 0) { ci = 0; frac = 0; }
         else if (ci >= cam->cmap_size - 1) { ci = cam->cmap_size - 2; frac = 1.0; }
         else frac = di - (double)ci;
         for (int k = 0; k < 4; k++)
            col[k] = dmap[4 * ci + k] * (1.0 - frac) + dmap[4 * (ci + 1) + k] * frac;
      } else {
         if (ci < 0) ci = 0;
         else if (ci >= cam->cmap_size - 1) ci = cam->cmap_size - 1;
         for (int k = 0; k < 4; k++) col[k] = dmap[4 * ci + k];
      }
      const double vis = p[3];
      for (int k = 0; k < 4; k++) o_bump(&b[k], vis == 1.0 ? col[k] : vis * col[k]);
      o_bump(&b[4], vis == 1.0 ? 255.0 : vis * 255.0);
   }
}

/* ------------------------------------------------------------------------------------ */
/* filters (filters.c)                                                                    */
/* ------------------------------------------------------------------------------------ */
static double o_sinc(double x) { x *= M_PI; return x != 0 ? sin(x) / x : 1.0; }
static double o_gauss(double x) { return exp(-2.0 * x * x) * sqrt(2.0 / M_PI); }

double o_spatial_kernel(int k, double x) { /* numeric dispatch of filters.c:172-202 */
   const double t = fabs(x);
   switch (k) {
   case 0: return o_gauss(x);
   case 1: return t < 1 ? (2 * t - 3) * t * t + 1 : 0;                       /* hermite */
   case 2: return (x > -0.5 && x <= 0.5) ? 1 : 0;                            /* box */
   case 3: return t < 1 ? 1 - t : 0;                                         /* triangle */
   case 4: return t < .5 ? .75 - t * t : (t < 1.5 ? .5 * (t - 1.5) * (t - 1.5) : 0); /* bell */
   case 5: return t < 1 ? .5 * t * t * t - t * t + 2.0 / 3 : (t < 2 ? (2 - t) * (2 - t) * (2 - t) / 6 : 0);
   case 6: { /* mitchell, B = C = 1/3 */
      const double B = 1.0 / 3, C = 1.0 / 3, tt = x * x;
      if (t < 1) return ((12 - 9 * B - 6 * C) * (t * tt) + (-18 + 12 * B + 6 * C) * tt + (6 - 2 * B)) / 6;
      if (t < 2) return ((-B - 6 * C) * (t * tt) + (6 * B + 30 * C) * tt + (-12 * B - 48 * C) * t + (8 * B + 24 * C)) / 6;
      return 0;
   }
   case 7: return o_sinc(x) * (0.42 + 0.5 * cos(M_PI * x) + 0.08 * cos(2 * M_PI * x));
   case 8: /* catrom */
      if (x < -2) return 0;
      if (x < -1) return 0.5 * (4 + x * (8 + x * (5 + x)));
      if (x < 0) return 0.5 * (2 + x * x * (-5 - 3 * x));
      if (x < 1) return 0.5 * (2 + x * x * (-5 + 3 * x));
      if (x < 2) return 0.5 * (4 + x * (-8 + x * (5 - x)));
      return 0;
   case 9: return o_sinc(x) * (0.5 + 0.5 * cos(M_PI * x));
   case 10: return o_sinc(x) * (0.54 + 0.46 * cos(M_PI * x));
   case 11: return (t < 3 ? o_sinc(t) * o_sinc(t / 3) : 0) * o_sinc(x / 3);
   case 12: return (t < 2 ? o_sinc(t) * o_sinc(t / 2) : 0) * o_sinc(x / 2);
   default: /* quadratic */
      if (x < -1.5) return 0;
      if (x < -0.5) return 0.5 * (x + 1.5) * (x + 1.5);
      if (x < 0.5) return 0.75 - x * x;
      if (x < 1.5) return 0.5 * (x - 1.5) * (x - 1.5);
      return 0;
   }
}

/* filters.c:217-270; out must hold at least 64*64 doubles; returns width or -1 */
int o_spatial_filter(int select, int oversample, double radius, double aspect, int field, double *out) {
   static const double support[14] = {1.5, 1, .5, 1, 1.5, 2, 2, 1, 2, 1, 1, 3, 2, 1.5};
   const double fw = 2.0 * support[select] * oversample * radius / aspect;
   int w = (int)fw + 1;
   if ((w ^ oversample) & 1) w++;
   const double adjust = fw > 0 ? support[select] * w / fw : 1.0;
   double sum = 0;
   for (int i = 0; i < w; i++)
      for (int j = 0; j < w; j++) {
         const double ii = ((2.0 * i + 1.0) / w - 1.0) * adjust;
         double jj = ((2.0 * j + 1.0) / w - 1.0) * adjust;
         if (field) jj *= 2.0;
         jj /= aspect;
         sum += out[i + j * w] = o_spatial_kernel(select, ii) * o_spatial_kernel(select, jj);
      }
   if (sum == 0.0) return -1;
   for (int i = 0; i < w * w; i++) out[i] *= 1.0 / sum;
   return w;
}

/* filters.c:409-489; returns sumfilt */
double o_temporal_filter(int n, int type, double fexp, double fwidth, double *weights, double *deltas) {
   if (n == 1) { deltas[0] = 0; weights[0] = 1; return 1.0; }
   double mx = 0, sum = 0;
   for (int i = 0; i < n; i++) {
      deltas[i] = ((double)i / (double)(n - 1) - 0.5) * fwidth;
      if (type == 2) weights[i] = pow(fexp >= 0 ? (i + 1.0) / n : (double)(n - i) / n, fabs(fexp));
      else if (type == 1) weights[i] = o_gauss(1.5 * fabs(i - n / 2.0) / (n / 2.0));
      else weights[i] = 1.0;
      if (weights[i] > mx) mx = weights[i];
   }
   for (int i = 0; i < n; i++) sum += weights[i] /= mx;
   return sum / n;
}

typedef struct {
   int max_filtered_counts, max_filter_index, kernel_size, n_kernels;
   double *widths, *coefs;
} o_de;

/* filters.c:272-407 */
int o_de_create(double max_rad, double min_rad, double curve, int ss, o_de *de) {
   memset(de, 0, sizeof(*de));
   de->kernel_size = -1;
   if (curve <= 0.0 || max_rad < min_rad) return 1;
   const double rmax = max_rad * ss + 1, rmin = min_rad * ss + 1;
   const double nd = pow(rmax / rmin, 1.0 / curve);
   if (nd > 1e7) return 1;
   const int nf = (int)ceil(nd);
   int rows;
   if (nf > 100) {
      rows = (int)ceil(100 + pow(nf - 100, curve)) + 1;
      de->max_filtered_counts = (int)pow((double)(rows - 100), 1.0 / curve) + 100;
   } else
      rows = de->max_filtered_counts = nf;
   const int half = ((int)(2 * ceil(rmax) - 1) - 1) / 2;
   de->kernel_size = (half + 1) * (2 + half) / 2;
   de->n_kernels = rows;
   de->coefs = (double *)calloc((size_t)rows * de->kernel_size, sizeof(double));
   de->widths = (double *)calloc((size_t)rows, sizeof(double));
   for (int f = 0; f < rows; f++) {
      double h = f < 100 ? rmax / pow(f + 1, curve) : rmax / pow(pow(f - 100, 1.0 / curve) + 100 + 1, curve);
      if (h <= rmin) { h = rmin; de->max_filter_index = f; }
      de->widths[f] = h;
      double sum = 0;
      for (int j = -half; j <= half; j++)
         for (int k = -half; k <= half; k++) {
            const double d = sqrt((double)(j * j + k * k)) / h;
            if (d <= 1.0) sum += o_gauss(1.5 * d);
         }
      int idx = f * de->kernel_size;
      for (int j = 0; j <= half; j++)
         for (int k = 0; k <= j; k++) {
            const double d = sqrt((double)(j * j + k * k)) / h;
            de->coefs[idx++] = d > 1.0 ? 0.0 : o_gauss(1.5 * d) / sum;
         }
      if (de->max_filter_index > 0) break;
   }
   if (de->max_filter_index == 0) de->max_filter_index = rows - 1;
   return 0;
}
void o_de_free(o_de *de) { free(de->widths); free(de->coefs); }

/* ------------------------------------------------------------------------------------ */
/* log scale + density estimation, scatter order of de_thread (rect.c:57-178), bits=33    */
/* ------------------------------------------------------------------------------------ */
static void o_acc(float *acc, int i, int ii, int j, int jj, int w, int h, const double *c) {
   if (j + jj >= 0 && j + jj < h && i + ii >= 0 && i + ii < w) {
      float *a = acc + 4 * ((size_t)(i + ii) + (size_t)(j + jj) * w);
      for (int k = 0; k < 4; k++) a[k] += c[k];
   }
}

void o_density(const uint32_t *buckets, float *accum, int w, int h, int os, double k1, double k2,
               double curve, const o_de *de) {
   if (!de || de->max_filter_index == 0) { /* rect.c:947-974 */
      for (size_t i = 0; i < (size_t)w * h; i++) {
         const uint32_t *b = buckets + 5 * i;
         if (b[3] == 0) continue;
         const double c3 = (double)b[3], ls = (k1 * log(1.0 + c3 * k2)) / c3;
         for (int k = 0; k < 4; k++) accum[4 * i + k] += (double)b[k] * ls;
      }
      return;
   }
   const int ss = (int)floor(os / 2.0), scf = !(os & 1);
   const double scfact = pow(os / (os + 1.0), 2.0);
   /* the reference's row loop runs one row past the array; rows and the
    * density neighbourhood are clipped to the array here */
   for (int j = os - 1; j <= h - os; j++)
      for (int i = os - 1; i < w - (os - 1); i++) {
         const uint32_t *b = buckets + 5 * ((size_t)i + (size_t)j * w);
         if (b[4] == 0 || b[3] == 0) continue;
         double f = 0.0;
         for (int ii = -ss; ii <= ss; ii++)
            for (int jj = -ss; jj <= ss; jj++) {
               const int x = i + ii, y = j + jj;
               if (x >= 0 && x < w && y >= 0 && y < h) f += buckets[5 * ((size_t)x + (size_t)y * w) + 4] / 255.0;
            }
         if (scf) f *= scfact;
         int k;
         if (f > de->max_filtered_counts) k = de->max_filter_index;
         else if (f <= 100) k = (int)ceil(f) - 1;
         else k = 100 + (int)floor(pow(f - 100, curve));
         if (k > de->max_filter_index) k = de->max_filter_index;
         int ci = k * de->kernel_size;
         const int fw = (int)ceil(de->widths[k]) - 1;
         for (int jj = 0; jj <= fw; jj++)
            for (int ii = 0; ii <= jj; ii++, ci++) {
               if (de->coefs[ci] == 0.0) continue;
               double c[4];
               const double c3 = (double)b[3];
               const double ls = de->coefs[ci] * (k1 * log(1.0 + c3 * k2)) / c3;
               for (int q = 0; q < 4; q++) c[q] = (double)b[q] * ls;
               if (jj == 0 && ii == 0) o_acc(accum, i, 0, j, 0, w, h, c);
               else if (ii == 0) {
                  o_acc(accum, i, jj, j, 0, w, h, c); o_acc(accum, i, -jj, j, 0, w, h, c);
                  o_acc(accum, i, 0, j, jj, w, h, c); o_acc(accum, i, 0, j, -jj, w, h, c);
               } else if (jj == ii) {
                  o_acc(accum, i, ii, j, jj, w, h, c); o_acc(accum, i, -ii, j, jj, w, h, c);
                  o_acc(accum, i, ii, j, -jj, w, h, c); o_acc(accum, i, -ii, j, -jj, w, h, c);
               } else {
                  o_acc(accum, i, ii, j, jj, w, h, c); o_acc(accum, i, -ii, j, jj, w, h, c);
                  o_acc(accum, i, ii, j, -jj, w, h, c); o_acc(accum, i, -ii, j, -jj, w, h, c);
                  o_acc(accum, i, jj, j, ii, w, h, c); o_acc(accum, i, -jj, j, ii, w, h, c);
                  o_acc(accum, i, jj, j, -ii, w, h, c); o_acc(accum, i, -jj, j, -ii, w, h, c);
               }
            }
      }
}

/* ------------------------------------------------------------------------------------ */
/* tone map (palettes.c:199-349) and final pass (rect.c:1067-1247)                        */
/* ------------------------------------------------------------------------------------ */
double o_calc_alpha(double d, double g, double lin) {
   if (!(d > 0)) return 0;
   if (d < lin) {
      const double frac = d / lin;
      return (1.0 - frac) * d * (pow(lin, g) / lin) + frac * pow(d, g);
   }
   return pow(d, g);
}

static void o_rgb2hsv(const double *rgb, double *hsv) {
   const double r = rgb[0], g = rgb[1], b = rgb[2];
   const double mx = fmax(r, fmax(g, b)), mn = fmin(r, fmin(g, b)), del = mx - mn;
   const double s = mx != 0 ? del / mx : 0;
   double hh = 0;
   if (s != 0) {
      const double rc = (mx - r) / del, gc = (mx - g) / del, bc = (mx - b) / del;
      if (r == mx) hh = bc - gc;
      else if (g == mx) hh = 2 + rc - bc;
      else if (b == mx) hh = 4 + gc - rc;
      if (hh < 0) hh += 6;
   }
   hsv[0] = hh; hsv[1] = s; hsv[2] = mx;
}

static void o_hsv2rgb(const double *hsv, double *rgb) {
   double hh = hsv[0];
   const double s = hsv[1], v = hsv[2];
   while (hh >= 6) hh -= 6;
   while (hh < 0) hh += 6;
   const int j = (int)floor(hh);
   const double f = hh - j, p = v * (1 - s), q = v * (1 - s * f), t = v * (1 - s * (1 - f));
   const double tab[6][3] = {{v, t, p}, {q, v, p}, {p, v, t}, {p, q, v}, {t, p, v}, {v, p, q}};
   const int k = (j >= 0 && j < 6) ? j : 0;
   rgb[0] = tab[k][0]; rgb[1] = tab[k][1]; rgb[2] = tab[k][2];
}

void o_calc_newrgb(const double *c, double ls, double highpow, double *out) {
   if (ls == 0.0 || (c[0] == 0 && c[1] == 0 && c[2] == 0)) { out[0] = out[1] = out[2] = 0; return; }
   double maxa = -1, maxc = 0;
   for (int i = 0; i < 3; i++) {
      const double a = ls * (c[i] / 255.0);
      if (a > maxa) { maxa = a; maxc = c[i] / 255.0; }
   }
   const double newls = 255.0 / maxc;
   if (maxa > 255 && highpow >= 0) {
      double hsv[3];
      for (int i = 0; i < 3; i++) out[i] = newls * (c[i] / 255.0) / 255.0;
      o_rgb2hsv(out, hsv);
      hsv[1] *= pow(newls / ls, highpow);
      o_hsv2rgb(hsv, out);
      for (int i = 0; i < 3; i++) out[i] *= 255.0;
   } else {
      double adj = -highpow;
      if (adj > 1) adj = 1;
      if (maxa <= 255) adj = 1.0;
      for (int i = 0; i < 3; i++) out[i] = ((1.0 - adj) * newls + adj * ls) * (c[i] / 255.0);
   }
}

typedef struct {
   int image_width, image_height, oversample, de_offset, filter_width;
   int earlyclip, nchan, transp, bytes_per_channel;
   double gamma, vibrancy, background[3], highlight_power, gam_lin_thresh;
} o_final_desc;

static void o_tonemap(double *t, const o_final_desc *d) {
   const double g = 1.0 / d->gamma;
   double alpha, ls;
   if (t[3] <= 0) { alpha = 0; ls = 0; }
   else {
      const double tmp = t[3] / 255.0;
      alpha = o_calc_alpha(tmp, g, d->gam_lin_thresh);
      ls = d->vibrancy * 256.0 * alpha / tmp;
      if (alpha < 0) alpha = 0;
      if (alpha > 1) alpha = 1;
   }
   double nrgb[3];
   o_calc_newrgb(t, ls, d->highlight_power, nrgb);
   for (int i = 0; i < 3; i++) {
      double a = nrgb[i] + (1.0 - d->vibrancy) * 256.0 * pow(t[i] / 255.0, g);
      if (d->nchan <= 3 || d->transp == 0) a += (1.0 - alpha) * d->background[i];
      else a = alpha > 0 ? a / alpha : 0;
      t[i] = a > 255 ? 255 : (a < 0 ? 0 : a);
   }
   t[3] = alpha;
}

/* accum is modified in place when earlyclip is set, as in the reference. `out` rows are
 * row_stride bytes apart. */
void o_final(float *accum, int w, int h, const double *filter, const o_final_desc *d, void *out,
             size_t row_stride) {
   if (d->earlyclip)
      for (size_t i = 0; i < (size_t)w * h; i++) {
         double t[4] = {accum[4 * i], accum[4 * i + 1], accum[4 * i + 2], accum[4 * i + 3]};
         o_tonemap(t, d);
         for (int k = 0; k < 4; k++) accum[4 * i + k] = (float)t[k];
      }
   const int fw = d->filter_width;
   for (int j = 0; j < d->image_height; j++)
      for (int i = 0; i < d->image_width; i++) {
         const int x = d->de_offset + i * d->oversample, y = d->de_offset + j * d->oversample;
         double t[4] = {0, 0, 0, 0};
         for (int ii = 0; ii < fw; ii++)
            for (int jj = 0; jj < fw; jj++) {
               const double k = filter[ii + jj * fw];
               const float *a = accum + 4 * ((size_t)(x + ii) + (size_t)(y + jj) * w);
               for (int q = 0; q < 4; q++) t[q] += k * a[q];
            }
         if (!d->earlyclip) o_tonemap(t, d);
         unsigned char *p8 = (unsigned char *)out + row_stride * j + (size_t)d->nchan * d->bytes_per_channel * i;
         unsigned short *p16 = (unsigned short *)p8;
         for (int q = 0; q < 3; q++) {
            double a = t[q] > 255 ? 255 : (t[q] < 0 ? 0 : t[q]);
            if (d->bytes_per_channel == 2) p16[q] = (unsigned short)(a * 256.0);
            else p8[q] = (unsigned char)a;
         }
         const double al = t[3] > 1 ? 1 : (t[3] < 0 ? 0 : t[3]);
         if (d->nchan > 3) {
            if (d->bytes_per_channel == 2) p16[3] = d->transp ? (unsigned short)(al * 65535) : 65535;
            else p8[3] = d->transp ? (unsigned char)(al * 255) : 255;
         }
      }
}

/* ------------------------------------------------------------------------------------ */
/* batch helpers for the tests (same per-point seeding as tests/harness/f3_harness.c)     */
/* ------------------------------------------------------------------------------------ */
/* in[i] = {x, y, color}; out[i] = {x, y, color, vis, bad} */
void o_apply_points(const o_genome *g, int xi, const double *in, int n, double *out,
                    const char *seed) {
   for (int i = 0; i < n; i++) {
      o_rand rc;
      const double p[4] = {in[3 * i], in[3 * i + 1], in[3 * i + 2], 1.0};
      double q[4];
      memset(&rc, 0, sizeof(rc));
      strncpy((char *)rc.randrsl, seed, 16 * sizeof(o_ub4));
      rc.randrsl[15] ^= (o_ub4)i;
      o_irandinit(&rc, 1);
      const int bad = o_apply_xform(&g->xform[xi], p, q, &rc);
      memcpy(out + 5 * i, q, 4 * sizeof(double));
      out[5 * i + 4] = bad;
   }
}

/* flam3_iterate from start point p0 with a context seeded from `seed` */
int o_iterate_seeded(const o_genome *g, int n, int fuse, const double *p0, double *samples,
                     const char *seed) {
   unsigned short *table = (unsigned short *)malloc(sizeof(unsigned short) * O_GRAIN *
                                                    (size_t)(o_num_std(g) + 1));
   o_rand rc;
   if (!o_chooser_table(g, table)) {
      free(table);
      return -1;
   }
   o_seed(&rc, seed);
   memcpy(samples, p0, 4 * sizeof(double));
   const int bad = o_iterate(g, n, fuse, samples, table, &rc);
   free(table);
   return bad;
}
