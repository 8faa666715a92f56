/*
 * f3b_pack.c — host packer: one interpolated flam3_genome -> one packed temporal sample
 * (layout: f3b_pack.h). Compiled against the reference's public header flam3.h, like any
 * libflam3 client; it is the only place (with f3b_render.c) that reads flam3 structs.
 *
 * It restates, for the packed representation:
 *   prepare_precalc_flags   variations.c:2016-2126  (active variation list, A/T/Y flags,
 *                                                    vis_adjusted, has_post, has_preblur)
 *   xform_precalc + *_precalc   variations.c:1942-2014
 *   flam3_create_xform_distrib / flam3_create_chaos_distrib   flam3.c:76-224
 *   the dmap and camera set-up of render_rectangle              rect.c:778-827
 */
#include <math.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "f3b_host.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define P_EPS 1e-10 /* EPS, private.h:47 */

static inline uint32_t f2u(double v) {
   float f = (float)v;
   uint32_t u;
   memcpy(&u, &f, 4);
   return u;
}

/* interpolation.c:22-29 */
static double vis_adjust(double opacity) {
   if (opacity == 0.0) return 0.0;
   return pow(10.0, -log(1.0 / opacity) / log(2.0));
}

/* which per-point precalcs a variation consumes (variations.c:2050-2115) */
static unsigned var_precalc_flags(int id) {
   switch (id) {
   case VAR_POLAR: case VAR_HANDKERCHIEF: case VAR_HEART: case VAR_DISC: case VAR_EX:
   case VAR_JULIA: case VAR_FAN: case VAR_FAN2: case VAR_DISC2: case VAR_POLAR2:
      return F3B_XF_ATAN_XY;
   case VAR_SPIRAL: case VAR_HYPERBOLIC: case VAR_DIAMOND: case VAR_POWER: case VAR_RINGS:
   case VAR_RINGS2:
      return F3B_XF_ANGLES;
   case VAR_BLOB:
      return F3B_XF_ATAN_XY | F3B_XF_ANGLES;
   case VAR_JULIAN: case VAR_JULIASCOPE: case VAR_RADIAL_BLUR: case VAR_NGON:
   case VAR_SUPER_SHAPE: case VAR_FLOWER: case VAR_CONIC: case VAR_CPOW: case VAR_ESCHER:
   case VAR_WEDGE: case VAR_WEDGE_JULIA: case VAR_WEDGE_SPH: case VAR_WHORL: case VAR_LOG:
      return F3B_XF_ATAN_YX;
   default:
      return 0;
   }
}

/* does the variation read precalc_sqrt (f->precalc_sqrt in variations.c)? */
static int var_needs_sqrt(int id) {
   switch (id) {
   case VAR_HORSESHOE: case VAR_POLAR: case VAR_HANDKERCHIEF: case VAR_HEART: case VAR_DISC:
   case VAR_SPIRAL: case VAR_HYPERBOLIC: case VAR_DIAMOND: case VAR_EX: case VAR_JULIA:
   case VAR_FISHEYE: case VAR_POWER: case VAR_RINGS: case VAR_FAN: case VAR_BLOB:
   case VAR_FAN2: case VAR_RINGS2: case VAR_EYEFISH: case VAR_RADIAL_BLUR: case VAR_BLADE:
   case VAR_SECANT2: case VAR_TWINTRIAN: case VAR_SUPER_SHAPE: case VAR_FLOWER:
   case VAR_CONIC: case VAR_PARABOLA: case VAR_SCRY: case VAR_WEDGE: case VAR_WEDGE_SPH:
   case VAR_WHORL:
      return 1;
   default:
      return 0;
   }
}

/* Packed parameters of variation `id`, in the order f3b_variations.cuh consumes them.
 * Returns the number of words written (== f3b_var_nparams[id]). */
static int pack_var_params(const flam3_xform *x, int id, double *p) {
   switch (id) {
   case VAR_WAVES: { /* waves_precalc, variations.c:1969 */
      const double dx = x->c[2][0], dy = x->c[2][1];
      p[0] = x->c[1][0];
      p[1] = 1.0 / (dx * dx + P_EPS);
      p[2] = x->c[1][1];
      p[3] = 1.0 / (dy * dy + P_EPS);
      return 4;
   }
   case VAR_POPCORN: p[0] = x->c[2][0]; p[1] = x->c[2][1]; return 2;
   case VAR_RINGS: p[0] = x->c[2][0] * x->c[2][0] + P_EPS; return 1;
   case VAR_FAN:
      p[0] = M_PI * (x->c[2][0] * x->c[2][0] + P_EPS);
      p[1] = x->c[2][1];
      p[2] = 0.5 * p[0];
      return 3;
   case VAR_BLOB:
      p[0] = x->blob_low; p[1] = x->blob_high - x->blob_low; p[2] = x->blob_waves;
      return 3;
   case VAR_PDJ: p[0] = x->pdj_a; p[1] = x->pdj_b; p[2] = x->pdj_c; p[3] = x->pdj_d; return 4;
   case VAR_FAN2:
      p[0] = M_PI * (x->fan2_x * x->fan2_x + P_EPS);
      p[1] = x->fan2_y;
      p[2] = 0.5 * p[0];
      return 3;
   case VAR_RINGS2: p[0] = x->rings2_val * x->rings2_val + P_EPS; return 1;
   case VAR_PERSPECTIVE: { /* perspective_precalc, variations.c:1942 */
      const double ang = x->perspective_angle * M_PI / 2.0;
      p[0] = x->perspective_dist;
      p[1] = sin(ang);
      p[2] = x->perspective_dist * cos(ang);
      return 3;
   }
   case VAR_JULIAN: /* juliaN_precalc, variations.c:1948 */
      p[0] = x->julian_power;
      p[1] = x->julian_dist / x->julian_power / 2.0;
      p[2] = fabs(x->julian_power);
      return 3;
   case VAR_JULIASCOPE: /* juliaScope_precalc, variations.c:1959 */
      p[0] = x->juliascope_power;
      p[1] = x->juliascope_dist / x->juliascope_power / 2.0;
      p[2] = fabs(x->juliascope_power);
      return 3;
   case VAR_RADIAL_BLUR: /* radial_blur_precalc, variations.c:1964 */
      p[0] = sin(x->radial_blur_angle * M_PI / 2.0);
      p[1] = cos(x->radial_blur_angle * M_PI / 2.0);
      return 2;
   case VAR_PIE: p[0] = x->pie_slices; p[1] = x->pie_rotation; p[2] = x->pie_thickness; return 3;
   case VAR_NGON:
      p[0] = x->ngon_power / 2.0;
      p[1] = 2.0 * M_PI / x->ngon_sides;
      p[2] = x->ngon_corners;
      p[3] = x->ngon_circle;
      return 4;
   case VAR_CURL: p[0] = x->curl_c1; p[1] = x->curl_c2; return 2;
   case VAR_RECTANGLES: p[0] = x->rectangles_x; p[1] = x->rectangles_y; return 2;
   case VAR_DISC2: { /* disc2_precalc, variations.c:1977 */
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
   case VAR_SUPER_SHAPE: /* supershape_precalc, variations.c:1999 */
      p[0] = x->super_shape_m / 4.0;
      p[1] = x->super_shape_n2;
      p[2] = x->super_shape_n3;
      p[3] = x->super_shape_rnd;
      p[4] = x->super_shape_holes;
      p[5] = -1.0 / x->super_shape_n1;
      return 6;
   case VAR_FLOWER: p[0] = x->flower_petals; p[1] = x->flower_holes; return 2;
   case VAR_CONIC: p[0] = x->conic_eccentricity; p[1] = x->conic_holes; return 2;
   case VAR_PARABOLA: p[0] = x->parabola_height; p[1] = x->parabola_width; return 2;
   case VAR_BENT2: p[0] = x->bent2_x; p[1] = x->bent2_y; return 2;
   case VAR_BIPOLAR: p[0] = -(M_PI / 2.0) * x->bipolar_shift; return 1;
   case VAR_CELL: p[0] = x->cell_size; p[1] = 1.0 / x->cell_size; return 2;
   case VAR_CPOW:
      p[0] = 2.0 * M_PI / x->cpow_power;
      p[1] = x->cpow_r / x->cpow_power;
      p[2] = x->cpow_i / x->cpow_power;
      p[3] = x->cpow_power;
      return 4;
   case VAR_CURVE: {
      double xl = x->curve_xlength * x->curve_xlength, yl = x->curve_ylength * x->curve_ylength;
      if (xl < 1e-20) xl = 1e-20;
      if (yl < 1e-20) yl = 1e-20;
      p[0] = x->curve_xamp;
      p[1] = x->curve_yamp;
      p[2] = 1.0 / xl;
      p[3] = 1.0 / yl;
      return 4;
   }
   case VAR_ESCHER:
      p[0] = 0.5 * (1.0 + cos(x->escher_beta));
      p[1] = 0.5 * sin(x->escher_beta);
      return 2;
   case VAR_LAZYSUSAN:
      p[0] = x->lazysusan_spin; p[1] = x->lazysusan_space; p[2] = x->lazysusan_twist;
      p[3] = x->lazysusan_x; p[4] = x->lazysusan_y;
      return 5;
   case VAR_MODULUS: p[0] = x->modulus_x; p[1] = x->modulus_y; return 2;
   case VAR_OSCILLOSCOPE:
      p[0] = 2.0 * M_PI * x->oscope_frequency;
      p[1] = x->oscope_amplitude;
      p[2] = x->oscope_separation;
      p[3] = x->oscope_damping;
      return 4;
   case VAR_POPCORN2: p[0] = x->popcorn2_x; p[1] = x->popcorn2_y; p[2] = x->popcorn2_c; return 3;
   case VAR_SEPARATION:
      p[0] = x->separation_x * x->separation_x;
      p[1] = x->separation_xinside;
      p[2] = x->separation_y * x->separation_y;
      p[3] = x->separation_yinside;
      return 4;
   case VAR_SPLIT: p[0] = x->split_xsize * M_PI; p[1] = x->split_ysize * M_PI; return 2;
   case VAR_SPLITS: p[0] = x->splits_x; p[1] = x->splits_y; return 2;
   case VAR_STRIPES: p[0] = 1.0 - x->stripes_space; p[1] = x->stripes_warp; return 2;
   case VAR_WEDGE:
      p[0] = x->wedge_angle; p[1] = x->wedge_hole; p[2] = x->wedge_count; p[3] = x->wedge_swirl;
      p[4] = 1.0 - x->wedge_angle * x->wedge_count * M_1_PI * 0.5;
      return 5;
   case VAR_WEDGE_JULIA: /* wedgeJulia_precalc, variations.c:1953 */
      p[0] = x->wedge_julia_dist / x->wedge_julia_power / 2.0;
      p[1] = fabs(x->wedge_julia_power);
      p[2] = x->wedge_julia_power;
      p[3] = x->wedge_julia_count;
      p[4] = 1.0 - x->wedge_julia_angle * x->wedge_julia_count * M_1_PI * 0.5;
      p[5] = x->wedge_julia_angle;
      return 6;
   case VAR_WEDGE_SPH:
      p[0] = x->wedge_sph_angle; p[1] = x->wedge_sph_hole; p[2] = x->wedge_sph_count;
      p[3] = x->wedge_sph_swirl;
      p[4] = 1.0 - x->wedge_sph_angle * x->wedge_sph_count * M_1_PI * 0.5;
      return 5;
   case VAR_WHORL: p[0] = x->whorl_inside; p[1] = x->whorl_outside; return 2;
   case VAR_WAVES2:
      p[0] = x->waves2_freqx; p[1] = x->waves2_scalex; p[2] = x->waves2_freqy;
      p[3] = x->waves2_scaley;
      return 4;
   case VAR_AUGER:
      p[0] = x->auger_sym; p[1] = x->auger_weight; p[2] = x->auger_freq; p[3] = x->auger_scale;
      return 4;
   case VAR_FLUX: p[0] = 2.0 + x->flux_spread; return 1;
   case VAR_MOBIUS:
      p[0] = x->mobius_re_a; p[1] = x->mobius_im_a; p[2] = x->mobius_re_b; p[3] = x->mobius_im_b;
      p[4] = x->mobius_re_c; p[5] = x->mobius_im_c; p[6] = x->mobius_re_d; p[7] = x->mobius_im_d;
      return 8;
   default:
      return 0;
   }
}

/* One row of the xform chooser as cumulative counts: cum[j] = number of entries of the
 * reference's 16384-entry table whose value is <= j. Walks the same double-precision
 * accumulation as flam3_create_chaos_distrib (flam3.c:165-224), so the step positions
 * are identical. xi < 0 -> plain weights; else weights scaled by chaos[xi][.] */
/* First index i in [0, n] with r_i >= T for every T of thr[0..nthr) (non-decreasing), where
 * r_0 = 0 and r_{i+1} = fl(r_i + dr) in double arithmetic, exactly as the reference's loop
 * accumulates it (flam3.c:197-218). The 16384 dependent additions cost ~20 us per row, too
 * much when a frame packs 1000 temporal samples, so the sequence is evaluated in closed form
 * binade by binade: while r stays in [2^k, 2^(k+1)) every r_i is a multiple of ulp = 2^(k-52)
 * and fl(r + dr) = r + rn(dr / ulp) * ulp, a constant step (round-to-nearest-even ties are
 * walked one by one; they need dr's lowest set bit to sit exactly half an ulp below, which
 * only one binade can have). Binade crossings are done with real additions.
 * tests/test_cpu_boundary.py checks it against the plain loop on random and adversarial weights. */
static void first_crossings(double dr, unsigned n, const double *thr, int nthr, uint32_t *first) {
   int j = 0;
   double r = 0.0;
   unsigned i = 0;
   while (j < nthr && r >= thr[j]) first[j++] = 0; /* thresholds <= 0 */
   while (j < nthr && i < n) {
      /* one real step: also how a binade is entered */
      r += dr;
      i++;
      while (j < nthr && r >= thr[j]) first[j++] = i;
      if (j >= nthr || i >= n) break;
      int e;
      (void)frexp(r, &e); /* r in [2^(e-1), 2^e) */
      const int sh = 53 - e; /* r * 2^sh is an integer below 2^53 */
      const double dq = ldexp(dr, sh);
      if (!(dq >= 1.0) || dq >= 9007199254740992.0) continue; /* step below an ulp or huge: walk */
      const double fl = floor(dq);
      const double frac = dq - fl;
      if (frac == 0.5) { /* tie binade: walk it with real additions */
         const double top = ldexp(1.0, e);
         while (j < nthr && i < n && r < top) {
            r += dr;
            i++;
            while (j < nthr && r >= thr[j]) first[j++] = i;
         }
         continue;
      }
      const uint64_t d_int = (uint64_t)fl + (frac > 0.5 ? 1u : 0u);
      uint64_t s_int = (uint64_t)ldexp(r, sh);
      const uint64_t lim = 1ull << 53;
      /* steps that stay inside the binade and inside the table */
      uint64_t m_max = (lim - 1 - s_int) / d_int;
      if (m_max > (uint64_t)(n - i)) m_max = n - i;
      while (j < nthr) {
         /* does thr[j] fall inside this linear stretch? */
         int te;
         (void)frexp(thr[j], &te);
         uint64_t m_need;
         if (thr[j] <= 0.0 || te < e) m_need = 0; /* already passed (cannot happen: handled above) */
         else if (te > e) break;                    /* beyond this binade */
         else {
            const uint64_t t_int = (uint64_t)ldexp(thr[j], sh);
            if (t_int <= s_int) m_need = 0;
            else m_need = (t_int - s_int + d_int - 1) / d_int;
         }
         if (m_need > m_max) break;
         first[j++] = i + (uint32_t)m_need;
      }
      if (m_max > 0) {
         s_int += m_max * d_int;
         i += (unsigned)m_max;
         r = ldexp((double)s_int, -sh);
      }
   }
   while (j < nthr) first[j++] = n; /* never reached inside the table */
}

static int chooser_row(const flam3_genome *cp, int nstd, int xi, uint32_t *cum) {
   double dr = 0.0;
   for (int i = 0; i < nstd; i++) {
      double d = cp->xform[i].density;
      if (xi >= 0) d *= cp->chaos[xi][i];
      if (d < 0.0) {
         fprintf(stderr, "xform weight must be non-negative, not %g.\n", d);
         return 1;
      }
      dr += d;
   }
   if (dr == 0.0) {
      fprintf(stderr, "cannot iterate empty flame.\n");
      return 1;
   }
   dr /= (double)F3B_CHOOSER_GRAIN;

   /* The reference's loop (flam3.c:197-218): table[i] = j where j advances while r >= t,
    * t being the running sum of the weights and r the running sum of dr. Entry i therefore
    * holds the number of running sums t_k <= r_i (capped at nstd-1), and the cumulative
    * count of xform k is the first i with r_i >= t_k. */
   double thr[nstd > 1 ? nstd : 1];
   double t = 0.0;
   for (int k = 0; k < nstd - 1; k++) {
      t += (xi >= 0) ? cp->xform[k].density * cp->chaos[xi][k] : cp->xform[k].density;
      thr[k] = t;
   }
   first_crossings(dr, F3B_CHOOSER_GRAIN, thr, nstd - 1, cum);
   cum[nstd - 1] = F3B_CHOOSER_GRAIN;
   return 0;
}

/* the plain loop, kept for the tests that pin the closed form above to it */
int f3b_debug_chooser_counts(double dr, unsigned n, const double *thr, int nthr, uint32_t *first,
                             int closed_form) {
   if (closed_form) {
      first_crossings(dr, n, thr, nthr, first);
      return 0;
   }
   int j = 0;
   double r = 0.0;
   for (unsigned i = 0; i <= n && j < nthr; i++) {
      while (j < nthr && r >= thr[j]) first[j++] = i;
      r += dr;
   }
   while (j < nthr) first[j++] = n;
   return 0;
}

int f3b_chaos_enabled(const flam3_genome *cp, int nstd) {
   /* flam3_check_unity_chaos, flam3.c:148-163 */
   for (int i = 0; i < nstd; i++)
      for (int j = 0; j < nstd; j++)
         if (fabs(cp->chaos[i][j] - 1.0) > P_EPS) return 1;
   return 0;
}

static int pack_xform(const flam3_xform *x, uint32_t *w, size_t cap, size_t *pos) {
   size_t o = (*pos + 3) & ~(size_t)3; /* records are 16-byte aligned */
   if (o + F3B_X_VARS > cap) return 1;
   uint32_t *r = w + o;
   memset(r, 0, F3B_X_VARS * 4);
   *pos = o;
   r[F3B_X_C00] = f2u(x->c[0][0]); r[F3B_X_C10] = f2u(x->c[1][0]); r[F3B_X_C20] = f2u(x->c[2][0]);
   r[F3B_X_C01] = f2u(x->c[0][1]); r[F3B_X_C11] = f2u(x->c[1][1]); r[F3B_X_C21] = f2u(x->c[2][1]);
   r[F3B_X_P00] = f2u(x->post[0][0]); r[F3B_X_P10] = f2u(x->post[1][0]); r[F3B_X_P20] = f2u(x->post[2][0]);
   r[F3B_X_P01] = f2u(x->post[0][1]); r[F3B_X_P11] = f2u(x->post[1][1]); r[F3B_X_P21] = f2u(x->post[2][1]);
   r[F3B_X_SCOLOR] = f2u(x->color_speed * x->color);
   r[F3B_X_ONEMS] = f2u(1.0 - x->color_speed);
   r[F3B_X_VIS] = f2u(vis_adjust(x->opacity));
   r[F3B_X_PREBLUR] = f2u(x->var[VAR_PRE_BLUR]);
   r[F3B_X_WLINEAR] = f2u(x->var[VAR_LINEAR]);
   r[F3B_X_WSPHERICAL] = f2u(x->var[VAR_SPHERICAL]);

   unsigned flags = 0;
   const int has_post = !(x->post[0][0] == 1.0 && x->post[0][1] == 0.0 && x->post[1][0] == 0.0 &&
                          x->post[1][1] == 1.0 && x->post[2][0] == 0.0 && x->post[2][1] == 0.0);
   if (has_post) flags |= F3B_XF_HAS_POST;
   if (x->var[VAR_PRE_BLUR] != 0.0) flags |= F3B_XF_PREBLUR;

   o += F3B_X_VARS;
   unsigned nv = 0;
   for (int id = 0; id < flam3_nvariations; id++) {
      if (x->var[id] == 0.0) continue;
      /* pre_blur is listed by the reference but has no case in its switch
       * (variations.c:2181-2379): it only acts through has_preblur */
      if (id == VAR_PRE_BLUR) continue;
      /* linear and spherical live in fixed slots of the record header */
      if (id == VAR_LINEAR || id == VAR_SPHERICAL) continue;
      double p[8];
      const int np = pack_var_params(x, id, p);
      if (np != f3b_var_nparams[id]) {
         fprintf(stderr, "flam3_b200: internal error, variation %d packs %d params\n", id, np);
         return 1;
      }
      const int len = (np + 1) & ~1; /* even length keeps records 8-byte aligned */
      if (o + 2 + (size_t)len > cap) return 1;
      w[o++] = (uint32_t)id | ((uint32_t)len << 8);
      w[o++] = f2u(x->var[id]);
      for (int k = 0; k < len; k++) w[o++] = k < np ? f2u(p[k]) : 0u;
      flags |= var_precalc_flags(id);
      if (var_needs_sqrt(id)) flags |= F3B_XF_NEED_SQRT;
      nv += 2 + (unsigned)len;
   }
   if (flags & F3B_XF_ANGLES) flags |= F3B_XF_NEED_SQRT; /* sina/cosa divide by it */
   r[F3B_X_FLAGS] = flags | (nv << 8); /* nv = total words of the variation list */
   *pos = o;
   return 0;
}

int f3b_camera_setup(const flam3_genome *cp, const f3b_frame_geom *g, f3b_camera *cam) {
   /* rect.c:785-827 */
   if (cp->sample_density <= 0.0) {
      fprintf(stderr, "sample density (quality) must be greater than zero, not %g.\n",
              cp->sample_density);
      return 1;
   }
   const double scale = pow(2.0, cp->zoom);
   cam->sample_density = cp->sample_density * scale * scale;
   double ppux = cp->pixels_per_unit * scale;
   const double ppuy = g->field ? (ppux / 2.0) : ppux;
   ppux /= g->pixel_aspect_ratio;
   double shift = 0.0;
   if (g->field == flam3_field_even) shift = -0.5;
   else if (g->field == flam3_field_odd) shift = 0.5;
   shift = shift / ppux;
   const double t0 = (double)g->gutter_width / (g->oversample * ppux);
   const double t1 = (double)g->gutter_width / (g->oversample * ppuy);
   const double corner0 = cp->center[0] - g->image_width / ppux / 2.0;
   const double corner1 = cp->center[1] - g->image_height / ppuy / 2.0;
   cam->bounds[0] = corner0 - t0;
   cam->bounds[1] = corner1 - t1 + shift;
   cam->bounds[2] = corner0 + g->image_width / ppux + t0;
   cam->bounds[3] = corner1 + g->image_height / ppuy + t1 + shift;
   const double size0 = 1.0 / (cam->bounds[2] - cam->bounds[0]);
   const double size1 = 1.0 / (cam->bounds[3] - cam->bounds[1]);
   cam->rot[0][0] = cos(cp->rotate * 2 * M_PI / 360.0);
   cam->rot[0][1] = -sin(cp->rotate * 2 * M_PI / 360.0);
   cam->rot[1][0] = -cam->rot[0][1];
   cam->rot[1][1] = cam->rot[0][0];
   cam->ws0 = g->width * size0;
   cam->hs1 = g->height * size1;
   cam->ppux = ppux;
   cam->ppuy = ppuy;
   return 0;
}

int f3b_pack_sample(const flam3_genome *cp, const f3b_camera *cam, double color_scalar,
                    int cmap_size, int floor_adds, uint32_t *w, size_t cap, size_t *used,
                    int *count_is_alpha) {
   const int has_final_slot = cp->final_xform_index >= 0;
   const int nstd = cp->num_xforms - has_final_slot;
   const int final_on = (cp->final_xform_enable == 1 && has_final_slot);
   if (nstd <= 0) {
      fprintf(stderr, "cannot iterate empty flame.\n");
      return 1;
   }
   for (int i = 0; i < cp->num_xforms; i++)
      if (cp->xform[i].density < 0.0) {
         fprintf(stderr, "xform %d weight must be non-negative, not %g.\n", i,
                 cp->xform[i].density);
         return 1;
      }
   const int xaos = f3b_chaos_enabled(cp, nstd);
   const int rows = xaos ? nstd + 1 : 1;
   const int nxf = nstd + (final_on ? 1 : 0);

   /* chooser rows are padded to a multiple of 4 entries with 16384 ("never picked") so the
    * device compares four thresholds per LDS.128 */
   const int stride = (nstd + 3) & ~3;
   size_t o = F3B_H_SIZE;
   if (o + (size_t)rows * stride + (size_t)nxf > cap) return 2;
   memset(w, 0, F3B_H_SIZE * 4);
   w[F3B_H_NSTD] = (uint32_t)nstd;
   w[F3B_H_FINAL] = final_on ? (uint32_t)nstd : (uint32_t)-1;
   w[F3B_H_XAOS] = (uint32_t)xaos;
   w[F3B_H_FINAL_OPACITY] = f2u(final_on ? cp->xform[cp->final_xform_index].opacity : 1.0);
   w[F3B_H_PALMODE] = (cp->palette_mode == flam3_palette_mode_linear) ? 1u : 0u;
   w[F3B_H_CMAPSIZE] = (uint32_t)cmap_size;
   w[F3B_H_ROTATE] = (cp->rotate != 0.0) ? 1u : 0u;
   w[F3B_H_ROT00] = f2u(cam->rot[0][0]); w[F3B_H_ROT01] = f2u(cam->rot[0][1]);
   w[F3B_H_ROT10] = f2u(cam->rot[1][0]); w[F3B_H_ROT11] = f2u(cam->rot[1][1]);
   w[F3B_H_RCX] = f2u(cp->rot_center[0]); w[F3B_H_RCY] = f2u(cp->rot_center[1]);
   w[F3B_H_B0] = f2u(cam->bounds[0]); w[F3B_H_B1] = f2u(cam->bounds[1]);
   w[F3B_H_B2] = f2u(cam->bounds[2]); w[F3B_H_B3] = f2u(cam->bounds[3]);
   w[F3B_H_WS0] = f2u(cam->ws0); w[F3B_H_HS1] = f2u(cam->hs1);
   w[F3B_H_NXF] = (uint32_t)nxf;

   /* chooser rows: row 0 plain weights, row i+1 conditioned on previous xform i */
   w[F3B_H_OFF_CHOOSER] = (uint32_t)o;
   w[F3B_H_CHOOSER_STRIDE] = (uint32_t)stride;
   for (int r = 0; r < rows; r++) {
      for (int j = nstd; j < stride; j++) w[o + j] = F3B_CHOOSER_GRAIN;
      if (chooser_row(cp, nstd, r - 1, w + o)) return 1;
      o += (size_t)stride;
   }

   /* xform table + records. The reference indexes standard xforms 0..nstd-1 and assumes
    * the final xform, if any, is the last one (flam3.c:176). */
   w[F3B_H_OFF_XFTAB] = (uint32_t)o;
   uint32_t *tab = w + o;
   o += (size_t)nxf;
   for (int i = 0; i < nstd; i++) {
      o = (o + 3) & ~(size_t)3;
      tab[i] = (uint32_t)o;
      if (pack_xform(&cp->xform[i], w, cap, &o)) return 2;
   }
   if (final_on) {
      o = (o + 3) & ~(size_t)3;
      tab[nstd] = (uint32_t)o;
      if (pack_xform(&cp->xform[cp->final_xform_index], w, cap, &o)) return 2;
   }
   /* Equal record stride: the device computes record k at base + k * stride instead of
    * looking the offset up (one dependent shared-memory load less per iteration). Records
    * were packed back to back; spread them out from the last one so nothing is clobbered. */
   {
      o = (o + 3) & ~(size_t)3;
      size_t xstride = 0;
      for (int i = 0; i < nxf; i++) {
         const size_t end = (i + 1 < nxf) ? tab[i + 1] : o;
         if (end - tab[i] > xstride) xstride = end - tab[i];
      }
      const size_t base = tab[0];
      if (base + xstride * (size_t)nxf > cap) return 2;
      for (int i = nxf - 1; i >= 0; i--) {
         const size_t end = (i + 1 < nxf) ? tab[i + 1] : o;
         const size_t len = end - tab[i], dst = base + xstride * (size_t)i;
         memmove(w + dst, w + tab[i], len * 4);
         memset(w + dst + len, 0, (xstride - len) * 4);
      }
      for (int i = 0; i < nxf; i++) tab[i] = (uint32_t)(base + xstride * (size_t)i);
      o = base + xstride * (size_t)nxf;
      w[F3B_H_XF_BASE] = (uint32_t)base;
      w[F3B_H_XF_STRIDE] = (uint32_t)xstride;
   }

   /* dmap (rect.c:778-782): palette * WHITE_LEVEL * temporal filter weight */
   o = (o + 3) & ~(size_t)3;
   if (o + 256 * 5 > cap) return 2;
   w[F3B_H_OFF_PALETTE] = (uint32_t)o;
   /* With step lookup, alpha == 1 and weight == 1 the alpha sum equals the count sum hit
    * by hit. Linear lookup blends alpha in floating point (rect.c:482-485), which the
    * truncating bucket add can round down, so it keeps its own count plane. */
   int alpha_ok = (cp->palette_mode != flam3_palette_mode_linear);
   /* all hits add dmap[i] unchanged when every xform is fully visible (vis_adjusted == 1
    * <=> opacity == 1) and the lookup is a plain step */
   int plot_ready = (cp->palette_mode != flam3_palette_mode_linear);
   for (int i = 0; i < cp->num_xforms; i++)
      if (cp->xform[i].opacity != 1.0) plot_ready = 0;
   int packed = plot_ready && floor_adds;
   const int truncate = plot_ready && floor_adds; /* what the integer bucket add keeps */
   for (int j = 0; j < 256; j++) {
      const double *col = cp->palette[j].color;
      uint32_t rgba = 0;
#if defined(__SSE2__)
      /* two channels at a time (this loop is 1024 conversions per temporal sample, 70 % of the
       * packer's time): the same multiplications in the same order; floor as a truncating
       * conversion, which is what it is for 0 <= c < 2^31 - anything else takes the scalar way */
      const __m128d k255 = _mm_set1_pd(255.0), ks = _mm_set1_pd(color_scalar);
      __m128d c01 = _mm_mul_pd(_mm_mul_pd(_mm_loadu_pd(col), k255), ks);
      __m128d c23 = _mm_mul_pd(_mm_mul_pd(_mm_loadu_pd(col + 2), k255), ks);
      const __m128d lo = _mm_setzero_pd(), hi = _mm_set1_pd(2147483648.0);
      const int in_range = _mm_movemask_pd(_mm_and_pd(_mm_cmpge_pd(c01, lo), _mm_cmplt_pd(c01, hi))) == 3 &&
                           _mm_movemask_pd(_mm_and_pd(_mm_cmpge_pd(c23, lo), _mm_cmplt_pd(c23, hi))) == 3;
      if (in_range) {
         const __m128i i01 = _mm_cvttpd_epi32(c01), i23 = _mm_cvttpd_epi32(c23);
         if (truncate) {
            c01 = _mm_cvtepi32_pd(i01);
            c23 = _mm_cvtepi32_pd(i23);
         }
         const __m128 f4 = _mm_movelh_ps(_mm_cvtpd_ps(c01), _mm_cvtpd_ps(c23));
         _mm_storeu_ps((float *)(w + o + j * 4), f4);
         const __m128i i4 = _mm_unpacklo_epi64(i01, i23); /* the four truncated values */
         /* all four are whole numbers <= 255 exactly when truncation changed nothing (or was
          * asked for) and no value exceeds 255 */
         const __m128d top = _mm_set1_pd(255.0);
         const int small = _mm_movemask_pd(_mm_cmple_pd(c01, top)) == 3 && _mm_movemask_pd(_mm_cmple_pd(c23, top)) == 3;
         if (small) {
            uint32_t v[4];
            _mm_storeu_si128((__m128i *)v, i4);
            rgba = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
         } else {
            packed = 0;
            uint32_t v[4];
            double c[4];
            _mm_storeu_si128((__m128i *)v, i4);
            _mm_storeu_pd(c, c01);
            _mm_storeu_pd(c + 2, c23);
            for (int k = 0; k < 4; k++)
               if (c[k] <= 255.0) rgba |= v[k] << (8 * k);
         }
      } else
#endif
      for (int k = 0; k < 4; k++) {
         double c = col[k] * 255.0 * color_scalar;
         if (truncate) c = floor(c);
         w[o + j * 4 + k] = f2u(c);
         if (!(c >= 0.0 && c <= 255.0)) packed = 0;
         else rgba |= (uint32_t)c << (8 * k);
      }
      w[o + 1024 + j] = rgba;
      if (col[3] * 255.0 * color_scalar != 255.0) alpha_ok = 0;
   }
   w[F3B_H_PLOT] = (plot_ready ? F3B_PLOT_READY : 0u) | (packed ? F3B_PLOT_PACKED : 0u);
   o += 256 * 5;
   w[F3B_H_WORDS] = (uint32_t)o;
   *used = o;
   if (count_is_alpha) *count_is_alpha = alpha_ok;
   return 0;
}
