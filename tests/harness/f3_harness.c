/*
 * tests/harness/f3_harness.c — TEST INFRASTRUCTURE. A few plain-C entry points over the
 * public flam3.h API so that Python (ctypes) can drive either the unmodified reference
 * library (oracle/_ref/libflam3_ref.so) or the B200 drop-in (dropin/_build/libflam3.so)
 * the same way flam3-render.c / flam3-animate.c do: parse genomes from XML, fill a
 * flam3_frame, seed it like flam3_init_frame (flam3.c:2494-2516), call flam3_render.
 * Compiled once against each library; contains no rendering logic of its own.
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "flam3.h"

typedef struct {
   int bits, bytes_per_channel, earlyclip, nthreads, sub_batch_size, verbose;
   int field, nchan, transp;
   int first_genome, ngenomes; /* slice of the parsed array handed to the frame */
   double time, pixel_aspect_ratio;
   char isaac_seed[128];
} h_render_args;

void *h_parse(const char *xml, int *ncps) {
   char *copy = strdup(xml);
   char name[] = "harness.flam3";
   flam3_genome *g = flam3_parse_xml2(copy, name, flam3_defaults_on, ncps);
   free(copy);
   return g;
}

void h_free_genomes(void *cps, int n) {
   flam3_genome *g = (flam3_genome *)cps;
   if (!g) return;
   for (int i = 0; i < n; i++) clear_cp(&g[i], 0);
   free(g);
}

int h_genome_dims(void *cps, int idx, int *w, int *h, int *oversample) {
   flam3_genome *g = (flam3_genome *)cps;
   *w = g[idx].width;
   *h = g[idx].height;
   *oversample = g[idx].spatial_oversample;
   return 0;
}

static void seed_frame(flam3_frame *f, const char *seed) {
   /* same steps as flam3_init_frame with isaac_seed set (flam3.c:2501-2515) */
   memset(f->rc.randrsl, 0, RANDSIZ * sizeof(ub4));
   strncpy((char *)&f->rc.randrsl, seed, RANDSIZ * sizeof(ub4));
   irandinit(&f->rc, 1);
}

/* stats_out = {badvals, num_iters, render_seconds}; rc_out (optional) receives the
 * frame's randctx after the call (sizeof(randctx) bytes). Returns flam3_render's value. */
int h_render(void *cps, const h_render_args *a, void *out, double *stats_out, void *rc_out) {
   flam3_genome *g = (flam3_genome *)cps;
   flam3_frame f;
   stat_struct st;
   memset(&f, 0, sizeof(f));
   memset(&st, 0, sizeof(st));
   f.pixel_aspect_ratio = a->pixel_aspect_ratio;
   f.genomes = g + a->first_genome;
   f.ngenomes = a->ngenomes;
   f.verbose = a->verbose;
   f.bits = a->bits;
   f.bytes_per_channel = a->bytes_per_channel;
   f.earlyclip = a->earlyclip;
   f.time = a->time;
   f.progress = 0;
   f.progress_parameter = 0;
   f.nthreads = a->nthreads;
   f.sub_batch_size = a->sub_batch_size;
   seed_frame(&f, a->isaac_seed);
   int rv = flam3_render(&f, out, a->field, a->nchan, a->transp, &st);
   if (stats_out) {
      stats_out[0] = st.badvals;
      stats_out[1] = (double)st.num_iters;
      stats_out[2] = (double)st.render_seconds;
   }
   if (rc_out) memcpy(rc_out, &f.rc, sizeof(randctx));
   return rv;
}

double h_memory_required(void *cps, int idx, int bits) {
   flam3_frame f;
   memset(&f, 0, sizeof(f));
   f.genomes = (flam3_genome *)cps + idx;
   f.ngenomes = 1;
   f.bits = bits;
   return flam3_render_memory_required(&f);
}

/* Geometry of the render buffers for a frame, by the formulas of rect.c:628-686, using
 * the library's own flam3_interpolate/flam3_create_spatial_filter. dims = {fic.width,
 * fic.height, filter_width, gutter_width, de_offset, oversample}. */
extern int flam3_create_spatial_filter(flam3_frame *spec, int field, double **filter);
int h_buffer_geometry(void *cps, const h_render_args *a, int *dims) {
   flam3_genome *g = (flam3_genome *)cps;
   flam3_frame f;
   flam3_genome cp;
   double *filter = NULL;
   memset(&f, 0, sizeof(f));
   memset(&cp, 0, sizeof(cp));
   f.pixel_aspect_ratio = a->pixel_aspect_ratio;
   f.genomes = g + a->first_genome;
   f.ngenomes = a->ngenomes;
   f.time = a->time;
   flam3_interpolate(f.genomes, f.ngenomes, f.time, 0, &cp);
   int os = cp.spatial_oversample;
   int iw = cp.width, ih = a->field ? cp.height / 2 : cp.height;
   int fw = flam3_create_spatial_filter(&f, a->field, &filter);
   free(filter);
   if (fw < 0) return 1;
   int gutter = (fw - os) / 2, maxde = 0, de_offset = 0;
   for (int i = 0; i < f.ngenomes; i++) {
      double e = f.genomes[i].estimator;
      int w = (int)ceil(e) * os;
      if (w > maxde) maxde = w;
   }
   if (maxde > 0) maxde += os - 1;
   if (maxde > gutter) {
      de_offset = maxde - gutter;
      gutter = maxde;
   }
   dims[0] = os * iw + 2 * gutter;
   dims[1] = os * ih + 2 * gutter;
   dims[2] = fw;
   dims[3] = gutter;
   dims[4] = de_offset;
   dims[5] = os;
   clear_cp(&cp, 0);
   return 0;
}

/* Print genome idx back to XML (explicit palette) — used to make self-contained fixtures. */
char *h_print_genome(void *cps, int idx) {
   return flam3_print_to_string((flam3_genome *)cps + idx);
}

/* sizes the Python side asserts against SURVEY.md §8(b) */
void h_abi_sizes(int *out) {
   out[0] = (int)sizeof(randctx);
   out[1] = (int)sizeof(flam3_xform);
   out[2] = (int)sizeof(flam3_genome);
   out[3] = (int)sizeof(flam3_frame);
}

/* ---- finer-grained probes used to pin the oracle and the packer ------------------- */

/* the xform chooser table of genome idx (flam3.c:76-111); out must hold rows*16384 */
int h_xform_distrib(void *cps, int idx, unsigned short *out, int *rows) {
   flam3_genome *g = (flam3_genome *)cps + idx;
   unsigned short *d = flam3_create_xform_distrib(g);
   if (!d) return 1;
   int nstd = g->num_xforms - (g->final_xform_index >= 0);
   int nr = g->chaos_enable ? nstd + 1 : 1;
   memcpy(out, d, (size_t)nr * 16384 * sizeof(unsigned short));
   *rows = nr;
   free(d);
   return 0;
}

/* apply_xform (variations.c:2129) of xform fn to n points in[i] = {x, y, color}; out[i] =
 * {x, y, color, vis, bad}. Every point gets a fresh ISAAC context seeded from `seed` and its
 * index, so variations that draw random numbers are reproducible per point. */
extern int prepare_precalc_flags(flam3_genome *);
extern void xform_precalc(flam3_genome *, int);
extern int apply_xform(flam3_genome *cp, int fn, double *p, double *q, randctx *rc);
int h_apply_xform(void *cps, int idx, int fn, const double *in, int n, double *out,
                  const char *seed) {
   flam3_genome *g = (flam3_genome *)cps + idx;
   if (prepare_precalc_flags(g)) return 1;
   for (int i = 0; i < g->num_xforms; i++) xform_precalc(g, i);
   for (int i = 0; i < n; i++) {
      randctx rc;
      double p[4] = {in[3 * i], in[3 * i + 1], in[3 * i + 2], 1.0}, q[4];
      memset(&rc, 0, sizeof(rc));
      strncpy((char *)rc.randrsl, seed, RANDSIZ * sizeof(ub4));
      rc.randrsl[RANDSIZ - 1] ^= (ub4)i;
      irandinit(&rc, 1);
      int bad = apply_xform(g, fn, p, q, &rc);
      out[5 * i] = q[0];
      out[5 * i + 1] = q[1];
      out[5 * i + 2] = q[2];
      out[5 * i + 3] = q[3];
      out[5 * i + 4] = bad;
   }
   return 0;
}

/* n draws of irand() (isaac.h:49-52) from a context seeded like flam3_init_frame; also
 * returns the context so another implementation can be run from the same state. */
int h_isaac_draws(const char *seed, int n, unsigned long *out, void *ctx_before, void *ctx_after) {
   randctx rc;
   memset(&rc, 0, sizeof(rc));
   strncpy((char *)rc.randrsl, seed, RANDSIZ * sizeof(ub4));
   irandinit(&rc, 1);
   if (ctx_before) memcpy(ctx_before, &rc, sizeof(rc));
   for (int i = 0; i < n; i++) out[i] = irand(&rc);
   if (ctx_after) memcpy(ctx_after, &rc, sizeof(rc));
   return 0;
}

/* flam3_iterate (flam3.c:232) on genome idx: n samples from start point p0[4]. */
int h_iterate(void *cps, int idx, int n, int fuse, const double *p0, double *samples,
              const char *seed) {
   flam3_genome *g = (flam3_genome *)cps + idx;
   randctx rc;
   if (prepare_precalc_flags(g)) return -1;
   unsigned short *d = flam3_create_xform_distrib(g);
   if (!d) return -1;
   memset(&rc, 0, sizeof(rc));
   strncpy((char *)rc.randrsl, seed, RANDSIZ * sizeof(ub4));
   irandinit(&rc, 1);
   memcpy(samples, p0, 4 * sizeof(double));
   int bad = flam3_iterate(g, n, fuse, samples, d, &rc);
   free(d);
   return bad;
}

/* the analysis helpers either side of the render (flam3.h:587-589), seeded like a frame */
int h_bounding_box(void *cps, int idx, double eps, int nsamples, const char *seed, double *out4) {
   flam3_frame f;
   memset(&f, 0, sizeof(f));
   seed_frame(&f, seed);
   double bmin[2] = {0, 0}, bmax[2] = {0, 0};
   int rv = flam3_estimate_bounding_box((flam3_genome *)cps + idx, eps, nsamples, bmin, bmax, &f.rc);
   out4[0] = bmin[0];
   out4[1] = bmin[1];
   out4[2] = bmax[0];
   out4[3] = bmax[1];
   return rv;
}
int h_colorhist(void *cps, int idx, int num_batches, const char *seed, double *hist256) {
   flam3_frame f;
   memset(&f, 0, sizeof(f));
   seed_frame(&f, seed);
   return flam3_colorhist((flam3_genome *)cps + idx, num_batches, &f.rc, hist256);
}

/* a progress callback like a GUI's (flam3.h:599-614, rect.c:324-373): counts its calls,
 * remembers the last percentage and ETA, asks for an abort at call number abort_at (0 = never) */
static struct { int calls, abort_at; double last_percent, last_eta, max_percent; } g_prog;
static int harness_progress(void *param, double percent, int stage, double eta) {
   (void)param;
   (void)stage;
   g_prog.calls++;
   g_prog.last_percent = percent;
   g_prog.last_eta = eta;
   if (percent > g_prog.max_percent) g_prog.max_percent = percent;
   return (g_prog.abort_at && g_prog.calls >= g_prog.abort_at) ? 1 : 0;
}
/* h_render with the callback installed; prog_out = {calls, max percent, last eta} */
int h_render_progress(void *cps, const h_render_args *a, void *out, double *stats_out, int abort_at,
                      double *prog_out) {
   flam3_genome *g = (flam3_genome *)cps;
   flam3_frame f;
   stat_struct st;
   memset(&f, 0, sizeof(f));
   memset(&st, 0, sizeof(st));
   memset(&g_prog, 0, sizeof(g_prog));
   g_prog.abort_at = abort_at;
   f.pixel_aspect_ratio = a->pixel_aspect_ratio;
   f.genomes = g + a->first_genome;
   f.ngenomes = a->ngenomes;
   f.verbose = a->verbose;
   f.bits = a->bits;
   f.bytes_per_channel = a->bytes_per_channel;
   f.earlyclip = a->earlyclip;
   f.time = a->time;
   f.progress = harness_progress;
   f.progress_parameter = 0;
   f.nthreads = a->nthreads;
   f.sub_batch_size = a->sub_batch_size;
   seed_frame(&f, a->isaac_seed);
   int rv = flam3_render(&f, out, a->field, a->nchan, a->transp, &st);
   if (stats_out) {
      stats_out[0] = st.badvals;
      stats_out[1] = (double)st.num_iters;
      stats_out[2] = (double)st.render_seconds;
   }
   prog_out[0] = g_prog.calls;
   prog_out[1] = g_prog.max_percent;
   prog_out[2] = g_prog.last_eta;
   return rv;
}

/* the repository's own front end (flam3_b200/csrc/f3b_frontend.c), present in the drop-in
 * library only: parse with it, so that the tests can print the result with the REFERENCE's
 * printer (h_print_genome) and compare with the reference's own parse */
extern flam3_genome *f3b_parse_flam3(const char *xml, int defaults_on, int *ncps) __attribute__((weak));
void *h_parse_own(const char *xml, int *ncps) {
   *ncps = 0;
   if (!f3b_parse_flam3) return NULL;
   return f3b_parse_flam3(xml, flam3_defaults_on, ncps);
}

/* byte ranges of flam3_genome that are not genome DATA (pointers, the name of the file it came
 * from): {offset, length} pairs, for tests that compare parsed genomes byte for byte */
int h_genome_nondata_ranges(int *out) {
   flam3_genome *g = 0;
   int n = 0;
#define RANGE(field) out[n++] = (int)(size_t)&g->field, out[n++] = (int)sizeof(g->field)
   RANGE(xform);
   RANGE(chaos);
   RANGE(edits);
   RANGE(parent_fname);
#undef RANGE
   return n / 2;
}
/* the xform array and one chaos row of genome idx, raw */
const void *h_genome_xforms(void *cps, int idx, int *num, int *final_index) {
   flam3_genome *g = (flam3_genome *)cps + idx;
   *num = g->num_xforms;
   *final_index = g->final_xform_index;
   return g->xform;
}
const double *h_genome_chaos_row(void *cps, int idx, int row) { return ((flam3_genome *)cps + idx)->chaos[row]; }
/* the same for flam3_xform: padding holes, values derived at render time (precalc flags and
 * values, the active-variation list) and the motion pointer */
int h_xform_nondata_ranges(int *out) {
   int n = 0;
#define SPAN(from, to) out[n++] = (int)(from), out[n++] = (int)((to) - (from))
   SPAN(offsetof(flam3_xform, vis_adjusted), offsetof(flam3_xform, padding));   /* variations.c:2034 */
   SPAN(offsetof(flam3_xform, padding) + sizeof(int), offsetof(flam3_xform, wind));
   SPAN(offsetof(flam3_xform, precalc_angles_flag), offsetof(flam3_xform, blob_low)); /* :2039-2102 */
   SPAN(offsetof(flam3_xform, persp_vsin), offsetof(flam3_xform, motion_freq));
   SPAN(offsetof(flam3_xform, motion_func) + sizeof(int), offsetof(flam3_xform, num_motion));
   SPAN(offsetof(flam3_xform, num_motion) + sizeof(int), sizeof(flam3_xform));
#undef SPAN
   return n / 2;
}
/* the motion elements of xform xf of genome idx (raw), *n = how many */
const void *h_xform_motion(void *cps, int idx, int xf, int *n) {
   flam3_genome *g = (flam3_genome *)cps + idx;
   *n = g->xform[xf].num_motion;
   return g->xform[xf].motion;
}
