/*
 * f3b_analysis.c — the two analysis helpers that sit either side of the render in
 * flam3-genome's pick loop, on the GPU:
 *
 *   flam3_estimate_bounding_box   flam3.h:588, flam3.c:3574-3639
 *   flam3_colorhist               flam3.h:587, flam3.c:351-394
 *
 * Both are `flam3_iterate` plus a little host arithmetic. The reference walks ONE chain of
 * 10^4-10^5 iterations, which a GPU runs at one iteration per ~0.4 us; here the same number of
 * samples comes from many independent chains (f3b_orbits), each started at a random point and
 * fused like the reference's (20 discarded iterations): the samples follow the same invariant
 * measure, the results agree with the reference's within its own sampling noise (tested), the
 * random numbers differ. Same signatures, same return values, same error paths; the frame's
 * ISAAC context is advanced by the two draws the reference spends on its start point.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "f3b_host.h"

struct f3b_ctx *f3b_render_acquire(void); /* a free GPU context of the drop-in ... */
void f3b_render_release(void);             /* ... and back */

static int pack_for_orbits(flam3_genome *cp, uint32_t **words, size_t *used) {
   f3b_camera cam;
   memset(&cam, 0, sizeof(cam));
   cam.ws0 = cam.hs1 = cam.ppux = cam.ppuy = 1.0;
   cam.rot[0][0] = cam.rot[1][1] = 1.0;
   cam.sample_density = 1.0;
   uint32_t *w = (uint32_t *)malloc(F3B_MAX_BLOB_WORDS * 4);
   if (!w) return 2;
   int alpha_ok = 0;
   const int rc = f3b_pack_sample(cp, &cam, 1.0, 256, 1, w, F3B_MAX_BLOB_WORDS, used, &alpha_ok);
   if (rc) {
      free(w);
      return rc;
   }
   *words = w;
   return 0;
}

static uint64_t seed_from(randctx *rc) {
   /* flam3_random_isaac_11 twice (the reference's start point): the same two draws */
   const uint64_t a = (uint64_t)f3b_isaac_next(rc) & 0xffffffffull;
   const uint64_t b = (uint64_t)f3b_isaac_next(rc) & 0xffffffffull;
   return (a << 32 | b) ^ 0x9e3779b97f4a7c15ull;
}

static int cmp_double(const void *a, const void *b) {
   const double x = *(const double *)a, y = *(const double *)b;
   return x < y ? -1 : (x > y ? 1 : 0);
}

int flam3_estimate_bounding_box(flam3_genome *cp, double eps, int nsamples, double *bmin, double *bmax,
                                randctx *rc) {
   if (nsamples <= 0) nsamples = 10000;
   uint32_t *words = NULL;
   size_t used = 0;
   if (pack_for_orbits(cp, &words, &used)) return -1; /* message printed, as by the reference */
   struct f3b_ctx *ctx = f3b_render_acquire();
   if (!ctx) {
      free(words);
      return -1;
   }
   /* chains of ~200 samples: enough of them to fill the GPU's latency, long enough that the
    * 20 fused iterations stay a small share */
   int n_chains = nsamples / 200;
   if (n_chains < 1) n_chains = 1;
   if (n_chains > 8192) n_chains = 8192;
   const int per = (nsamples + n_chains - 1) / n_chains;
   float *pts = (float *)malloc((size_t)n_chains * per * 4 * sizeof(float));
   double *xs = (double *)malloc(sizeof(double) * 2 * (size_t)nsamples);
   uint64_t bad = 0;
   int rv = -1;
   if (pts && xs && !f3b_orbits(ctx, words, used * 4, n_chains, per, 20, seed_from(rc), NULL, pts, &bad)) {
      double *ys = xs + nsamples;
      for (int i = 0; i < nsamples; i++) { /* the first nsamples of n_chains * per */
         xs[i] = pts[4 * (size_t)i];
         ys[i] = pts[4 * (size_t)i + 1];
      }
      const int bv = (int)((double)bad * nsamples / ((double)n_chains * per));
      if (bv / (double)nsamples > eps) eps = 3 * bv / (double)nsamples;
      if (eps > 0.3) eps = 0.3;
      const int low_target = (int)(nsamples * eps);
      const int high_target = nsamples - low_target;
      qsort(xs, (size_t)nsamples, sizeof(double), cmp_double);
      qsort(ys, (size_t)nsamples, sizeof(double), cmp_double);
      if (low_target == 0) {
         bmin[0] = xs[0];
         bmin[1] = ys[0];
         bmax[0] = xs[nsamples - 1];
         bmax[1] = ys[nsamples - 1];
      } else {
         bmin[0] = xs[low_target];
         bmax[0] = xs[high_target];
         bmin[1] = ys[low_target];
         bmax[1] = ys[high_target];
      }
      rv = bv;
   } else if (pts && xs)
      fprintf(stderr, "flam3_estimate_bounding_box (B200): %s\n", f3b_last_error(ctx));
   f3b_render_release();
   free(pts);
   free(xs);
   free(words);
   return rv;
}

int flam3_colorhist(flam3_genome *cp, int num_batches, randctx *rc, double *hist) {
   const int sbs = 10000;
   memset(hist, 0, 256 * sizeof(double));
   if (num_batches <= 0) return 0;
   uint32_t *words = NULL;
   size_t used = 0;
   if (pack_for_orbits(cp, &words, &used)) return 1;
   struct f3b_ctx *ctx = f3b_render_acquire();
   if (!ctx) {
      free(words);
      return 1;
   }
   /* the reference's batches are independent chains already (random start, 20 fused, 10000
    * stored): each becomes 50 chains of 200 */
   const int per = 200, n_chains = num_batches * (sbs / per);
   float *pts = (float *)malloc((size_t)n_chains * per * 4 * sizeof(float));
   int rv = 1;
   if (pts && !f3b_orbits(ctx, words, used * 4, n_chains, per, 20, seed_from(rc), NULL, pts, NULL)) {
      for (size_t i = 0; i < (size_t)n_chains * per; i++) {
         int c = (int)(pts[4 * i + 2] * 256);
         if (c < 0) c = 0;
         if (c > 255) c = 255;
         hist[c] += 1;
      }
      for (int i = 0; i < 256; i++) hist[i] /= (float)(num_batches * sbs);
      rv = 0;
   } else if (pts)
      fprintf(stderr, "flam3_colorhist (B200): %s\n", f3b_last_error(ctx));
   f3b_render_release();
   free(pts);
   free(words);
   return rv;
}
