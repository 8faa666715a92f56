/*
 * f3b_render.c — the drop-in boundary: `flam3_render` and `flam3_render_memory_required`
 * with the reference's exact C ABI (flam3.h:623,625), so flam3-render, flam3-animate and
 * flam3-genome link against it unchanged.
 *
 * This is the host half of render_rectangle (rect.c:529-1286) re-expressed as a driver of
 * the GPU stages in include/flam3_b200.h: it keeps the batch / temporal-sample structure,
 * the filter and gutter sizing, the camera, k1/k2, the averaging of gamma / vibrancy /
 * background, the bits / field / nchan / transp / earlyclip semantics, the stats and the
 * evolution of the frame's ISAAC context; the iter_thread / de_thread pthreads are gone.
 * All temporal samples of a batch are packed first and run in ONE chaos-game launch.
 *
 * Host-side genome interpolation (`flam3_interpolate`) is not part of the hot path and is
 * resolved from the libflam3 this file is dropped into.
 *
 * There is no CPU fallback: if no sm_100-class GPU is usable the call fails loudly.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "f3b_host.h"

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

/* The GPUs of this process. F3B_DEVICES ("0-7", "0,2,4", ...) spreads every frame over several
 * GPUs inside this one call: the unmodified flam3-render / flam3-animate then use the whole
 * node. Default: the single device F3B_DEVICE (0). An entry may repeat ("0,0": two contexts
 * on one GPU, which is how the multi-GPU path is tested on a one-GPU box). */
#define F3B_MAX_LOCAL 16
typedef struct {
   int n;
   f3b_ctx *ctx[F3B_MAX_LOCAL];
} f3b_group_t;

/* Independent frames may be rendered concurrently from different host threads (SURVEY §8(b)
 * Threading; README.txt:336-338: GUIs call flam3_render from worker threads). A SLOT is
 * everything one flam3_render call needs for itself: its GPU context(s), the page-locked batch
 * buffer, the cached DE filter bank. F3B_CONTEXTS=n (default 1, at most 8) slots exist; a call
 * takes a free one and waits when all are busy. g_lock guards the slot table and the
 * configuration knobs below, which must be set before rendering starts, not during. */
typedef struct {
   uint32_t *words;
   size_t cap, used; /* in words */
} blob_buf;
typedef struct {
   int valid;
   double max_rad, min_rad, curve;
   int oversample;
   f3b_de_filters de;
} de_cache_t;
typedef struct {
   f3b_group_t group;
   de_cache_t de;
   blob_buf blobs;
   double t0; /* start of the frame being rendered */
   int busy;
} f3b_slot;
#define F3B_MAX_SLOTS 8
static f3b_slot g_slots[F3B_MAX_SLOTS];
static int g_nslots = 0;
static pthread_cond_t g_slot_free = PTHREAD_COND_INITIALIZER;
static __thread f3b_slot *t_slot = NULL; /* the slot the calling thread holds */
#define g_group (t_slot->group)
#define g_de_cache (t_slot->de)
#define g_blobs (t_slot->blobs)
#define g_progress_t0 (t_slot->t0)

static int parse_devices(const char *e, int *dev, int cap) {
   int n = 0;
   while (*e && n < cap) {
      char *end;
      long a = strtol(e, &end, 10);
      if (end == e) break;
      long b = a;
      if (*end == '-') {
         const char *q = end + 1;
         b = strtol(q, &end, 10);
         if (end == q) b = a;
      }
      for (long d = a; d <= b && n < cap; d++) dev[n++] = (int)d;
      e = end;
      while (*e == ',' || *e == ' ') e++;
   }
   return n;
}

static f3b_ctx *get_ctx(void) {
   if (!g_group.n) {
      int dev[F3B_MAX_LOCAL], n = 0;
      const char *e = getenv("F3B_DEVICES");
      if (e) n = parse_devices(e, dev, F3B_MAX_LOCAL);
      if (n == 0) {
         e = getenv("F3B_DEVICE");
         dev[0] = e ? atoi(e) : 0;
         n = 1;
      }
      for (int i = 0; i < n; i++) {
         if (f3b_ctx_create(dev[i], &g_group.ctx[i])) {
            fprintf(stderr, "flam3_render (B200): cannot create a GPU context on device %d; "
                            "this build has no CPU path.\n", dev[i]);
            for (int j = 0; j < i; j++) f3b_ctx_destroy(g_group.ctx[j]);
            return NULL;
         }
      }
      for (int i = 0; i < n; i++)
         for (int j = 0; j < n; j++)
            if (i != j && f3b_peer_enable(g_group.ctx[i], g_group.ctx[j])) {
               fprintf(stderr, "flam3_render (B200): %s\n", f3b_last_error(g_group.ctx[i]));
               for (int k = 0; k < n; k++) f3b_ctx_destroy(g_group.ctx[k]);
               return NULL;
            }
      g_group.n = n;
   }
   return g_group.ctx[0];
}

/* take a free slot (creating its contexts on first use); NULL when no GPU is usable */
static f3b_ctx *slot_acquire(void) {
   pthread_mutex_lock(&g_lock);
   if (!g_nslots) {
      const char *e = getenv("F3B_CONTEXTS");
      g_nslots = e ? atoi(e) : 1;
      if (g_nslots < 1) g_nslots = 1;
      if (g_nslots > F3B_MAX_SLOTS) g_nslots = F3B_MAX_SLOTS;
   }
   f3b_slot *sl = NULL;
   for (;;) {
      for (int i = 0; i < g_nslots && !sl; i++)
         if (!g_slots[i].busy) sl = &g_slots[i];
      if (sl) break;
      pthread_cond_wait(&g_slot_free, &g_lock);
   }
   sl->busy = 1;
   t_slot = sl;
   f3b_ctx *c = get_ctx(); /* context creation is serialised by g_lock */
   if (!c) {
      sl->busy = 0;
      t_slot = NULL;
      pthread_cond_signal(&g_slot_free);
   }
   pthread_mutex_unlock(&g_lock);
   return c;
}
static void slot_release(void) {
   pthread_mutex_lock(&g_lock);
   if (t_slot) t_slot->busy = 0;
   t_slot = NULL;
   pthread_cond_signal(&g_slot_free);
   pthread_mutex_unlock(&g_lock);
}
/* for the other entry points of the drop-in that use a GPU context (f3b_analysis.c) */
f3b_ctx *f3b_render_acquire(void) { return slot_acquire(); }
void f3b_render_release(void) { slot_release(); }

/* Exposed so hosts embedding the library (and the tests / bench) can reach the context
 * the drop-in renders with (the first slot's), e.g. for per-stage GPU timings. Not to be used
 * while another thread renders. */
f3b_ctx *f3b_render_context_at(int i);
f3b_ctx *f3b_render_context(void) { return f3b_render_context_at(0); }
/* the i-th GPU context of the first slot (F3B_DEVICES), NULL beyond the last */
f3b_ctx *f3b_render_context_at(int i) {
   pthread_mutex_lock(&g_lock);
   if (!g_nslots) g_nslots = 1;
   f3b_slot *saved = t_slot;
   t_slot = &g_slots[0];
   f3b_ctx *c = get_ctx();
   c = (c && i >= 0 && i < g_group.n) ? g_group.ctx[i] : NULL;
   t_slot = saved;
   pthread_mutex_unlock(&g_lock);
   return c;
}

/* Render the next frame(s) from a caller-supplied histogram instead of running the chaos
 * game: `buckets5` is the reference's bits=33 bucket layout, uint32[5] per bucket
 * (flam3.c:3824), nbuckets must equal the frame's bucket grid. Pass NULL to return to
 * normal operation. Used to drive the post-pass from a histogram produced elsewhere
 * (another GPU, or the reference itself in the parity tests). */
static const uint32_t *g_inject = NULL;
static size_t g_inject_n = 0;
static int g_inject_preload = 0;
void f3b_render_inject_histogram(const uint32_t *buckets5, size_t nbuckets) {
   pthread_mutex_lock(&g_lock);
   g_inject = buckets5;
   g_inject_n = nbuckets;
   g_inject_preload = 0;
   pthread_mutex_unlock(&g_lock);
}
/* Same, but the chaos game still runs and adds on top of the supplied buckets (tests of the
 * saturating integer adds start from buckets close to UINT_MAX). */
void f3b_render_preload_histogram(const uint32_t *buckets5, size_t nbuckets) {
   pthread_mutex_lock(&g_lock);
   g_inject = buckets5;
   g_inject_n = nbuckets;
   g_inject_preload = buckets5 != NULL;
   pthread_mutex_unlock(&g_lock);
}

/* Knobs for hosts that embed the library (bench.py, the multi-GPU driver). Defaults give
 * the plain single-GPU drop-in behaviour.
 *  - device_only: leave the finished image on the GPU (no device->host copy); `out` is
 *    not written.
 *  - shard (rank, nranks): this process runs only its share of the chaos-game chains
 *    (f3b_iterate's rank/nranks) into its private histogram.
 *  - exchange callback: called after the chaos game of every batch, with the stream
 *    idle, so the host can sum the per-GPU histograms (e.g. an NCCL all-reduce on
 *    f3b_hist_color_ptr / f3b_hist_count_ptr) before density estimation runs. */
static int g_device_only = 0;
static int g_rank = 0, g_nranks = 1;
static void (*g_exchange)(f3b_ctx *, void *) = NULL;
static void *g_exchange_user = NULL;
/*  - strip gather: see include/flam3_b200_dropin.h (row-strip post-pass across ranks) */
typedef void (*f3b_gather_fn)(f3b_ctx *, void *image, size_t row_bytes, int rows_per, void *user);
static f3b_gather_fn g_gather = NULL;
static int g_stream_ordered = 0;
static void *g_gather_user = NULL;
/*  - peers (ranks in separate processes, one GPU each): where every rank's buckets and rank
 *    0's image are, as pointers valid on this process's device (CUDA IPC), and a callback that
 *    enqueues a barrier over all ranks on the render stream. With them set, rank r sums its
 *    strip of the frame out of the peers' memory and writes its image rows into rank 0's
 *    image (include/flam3_b200.h, "several GPUs on one frame"). */
static f3b_peer_hist g_peers[16];
static int g_npeers = 0;
static void *g_peer_image0 = NULL;
static void *g_peer_own_color = NULL; /* what this rank exported: checked against reallocation */
static void (*g_barrier)(f3b_ctx *, void *) = NULL;
static void *g_barrier_user = NULL;
void f3b_render_set_peers(int nranks, const f3b_peer_hist *peers, void *image0) {
   pthread_mutex_lock(&g_lock);
   g_npeers = 0;
   g_peer_image0 = NULL;
   g_peer_own_color = NULL;
   if (peers && nranks > 1 && nranks <= 16) {
      for (int i = 0; i < nranks; i++) g_peers[i] = peers[i];
      g_npeers = nranks;
      g_peer_image0 = image0;
   }
   pthread_mutex_unlock(&g_lock);
}
void f3b_render_set_barrier(void (*cb)(f3b_ctx *, void *), void *user) {
   pthread_mutex_lock(&g_lock);
   g_barrier = cb;
   g_barrier_user = user;
   pthread_mutex_unlock(&g_lock);
}
void f3b_render_set_device_only(int on) { g_device_only = on; }
void f3b_render_set_shard(int rank, int nranks) {
   if (nranks < 1 || rank < 0 || rank >= nranks) {
      rank = 0;
      nranks = 1;
   }
   g_rank = rank;
   g_nranks = nranks;
}
void f3b_render_set_stream_ordered(int on) { g_stream_ordered = on; }
void f3b_render_set_strip_gather(f3b_gather_fn cb, void *user) {
   g_gather = cb;
   g_gather_user = user;
}
void f3b_render_set_exchange(void (*cb)(f3b_ctx *, void *), void *user) {
   g_exchange = cb;
   g_exchange_user = user;
}

double flam3_render_memory_required(flam3_frame *spec) {
   /* flam3.c:3874-3887: the reference's planning model (bytes), kept verbatim in meaning so
    * that flam3-render's nstrips / use_mem logic (flam3-render.c:36-101) behaves the same */
   flam3_genome *cps = spec->genomes;
   int real_bits = spec->bits;
   if (33 == real_bits) real_bits = 32;
   return (double)cps[0].spatial_oversample * cps[0].spatial_oversample * (double)cps[0].width *
          cps[0].height * (real_bits / 8) * 9.0;
}

/* F3B_TIMING=1: host-side phase times of every flam3_render call on stderr */
static double now_ms(void) {
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* The DE filter bank depends only on (estimator, estimator_minimum, estimator_curve,
 * oversample): the last one built is kept (under g_lock), since animations and repeated
 * renders ask for the same bank every batch and building it costs ~0.15 ms of idle GPU. */

static int de_filters_cached(double max_rad, double min_rad, double curve, int oversample,
                             f3b_de_filters *out) {
   if (!(g_de_cache.valid && g_de_cache.max_rad == max_rad && g_de_cache.min_rad == min_rad &&
         g_de_cache.curve == curve && g_de_cache.oversample == oversample)) {
      if (g_de_cache.valid) f3b_de_filters_free(&g_de_cache.de);
      g_de_cache.valid = 0;
      if (f3b_de_filters_create(max_rad, min_rad, curve, oversample, &g_de_cache.de)) return 1;
      g_de_cache.valid = 1;
      g_de_cache.max_rad = max_rad;
      g_de_cache.min_rad = min_rad;
      g_de_cache.curve = curve;
      g_de_cache.oversample = oversample;
   }
   *out = g_de_cache.de; /* borrowed: owned by the cache */
   return 0;
}

static int poll_progress(flam3_frame *spec, double percent, int stage) {
   /* 0 continue, 1 abort, 2 pause (rect.c:324-373); the 4th argument is the estimated time
    * to completion in seconds, like rect.c:338-343 */
   if (!spec->progress) return 0;
   double eta = 0.0;
   if (percent > 0.0 && g_progress_t0 > 0.0) eta = (now_ms() - g_progress_t0) * 1e-3 * (100.0 - percent) / percent;
   int rv = (*spec->progress)(spec->progress_parameter, percent, stage, eta);
   while (rv == 2) {
      struct timespec ts = {0, 100000000};
      nanosleep(&ts, NULL);
      rv = (*spec->progress)(spec->progress_parameter, percent, stage, eta);
   }
   return rv == 1;
}


/* The buffer the packed samples of a batch are assembled in lives across frames (under
 * g_lock) and is page-locked: fresh pageable memory costs a page fault per 4 KB and a
 * staging copy per upload, which for an animation frame (1000 samples, 4.6 MB) was a
 * quarter of the host time. f3b_iterate's upload from it is asynchronous, so
 * f3b_inputs_consumed must be awaited before the buffer is written again. */

static int blob_reserve(blob_buf *b, size_t extra) {
   if (b->used + extra <= b->cap) return 0;
   size_t ncap = b->cap ? b->cap * 2 : 1u << 18;
   while (ncap < b->used + extra) ncap *= 2;
   uint32_t *n = (uint32_t *)f3b_pinned_alloc(ncap * 4);
   if (!n) return 1;
   if (b->used) memcpy(n, b->words, b->used * 4);
   f3b_pinned_free(b->words);
   b->words = n;
   b->cap = ncap;
   return 0;
}

/* ---- host-parallel packing of the temporal samples of one batch ------------------- */
typedef struct {
   int rc, alpha_ok; /* rc: 0 ok, 1 genome error, 2 too large, 3 camera error */
   size_t used;      /* words */
   uint32_t *words;  /* the packed sample: in its slot of the batch buffer, or malloc'd */
   int owned;        /* words was malloc'd (the sample did not fit its slot) */
   f3b_camera cam;
   double vibrancy, gamma, background[3];
   double contrast, brightness, estimator_curve, gam_lin_thresh;
} packed_sample;

typedef struct {
   packed_sample *s;
   int n;
   size_t slot_words; /* samples that are not `owned` sit at ts * slot_words of the batch buffer */
} packed_batch;

typedef struct {
   flam3_frame *spec;
   const f3b_frame_geom *g;
   const double *tfilter, *tdeltas;
   int batch_num, nts, cmap_size, floor_adds;
   volatile int next;
   packed_sample *out;
   uint32_t *slots;   /* batch buffer: sample ts is packed at slots + ts * slot_words ... */
   size_t slot_words; /* ... when it fits (0: no slots yet, sample 0 sizes them) */
} pack_job;

static void packed_free(packed_batch *b) {
   if (!b->s) return;
   for (int i = 0; i < b->n; i++)
      if (b->s[i].owned) free(b->s[i].words);
   free(b->s);
   b->s = NULL;
   b->n = 0;
   b->slot_words = 0;
}

/* ---- the genome of one temporal sample ---------------------------------------------------
 * flam3_interpolate (flam3.c:717-802) for a time between two key frames: make aligned copies
 * of the two genomes (flam3_align: pad to the same xforms), reset the result, and blend
 * everything with flam3_interpolate_n (interpolation.c:373-708). An animation frame asks for
 * 1000 such genomes, and that was its bottleneck: ~25 us each, 80 % of the host's work per
 * temporal sample, against ~3000 iterations on the GPU. Two thirds of it does not depend on
 * the time and is done once per worker and segment here:
 *   - the aligned copies (5.6 us);
 *   - the palettes' way into HSV: interpolate_n converts all 256 entries of both key frames
 *     (three rgb2hsv per entry) for every sample. The copies are switched to the "sweep"
 *     palette mode, which makes interpolate_n copy entries instead, and the blend is done
 *     here from HSV values kept with the copies - the same operations on the same numbers,
 *     with the reference's own rgb2hsv / hsv2rgb (bit-identical, tests/test_cpu_boundary.py).
 * Smooth (Catmull-Rom) segments and single genomes go through flam3_interpolate as before.
 * The symbols are weak: a program linked without libflam3's interpolation.c (tools/f3b-render)
 * renders single key frames only and never gets here. */
extern void flam3_align(flam3_genome *dst, flam3_genome *src, int nsrc) __attribute__((weak));
extern void flam3_interpolate_n(flam3_genome *result, int ncp, flam3_genome *cpi, double *c, double stagger)
    __attribute__((weak));
extern void flam3_add_xforms(flam3_genome *cp, int num_to_add, int interp_padding, int final_flag)
    __attribute__((weak));
extern void rgb2hsv(double *rgb, double *hsv) __attribute__((weak));
extern void hsv2rgb(double *hsv, double *rgb) __attribute__((weak));

typedef struct {
   int i1;       /* the copies are cps[i1], cps[i1 + 1] aligned; -1: none yet */
   int pal_mode; /* palette_interpolation of the first one (the copy says sweep) */
   flam3_genome cpi[2];
   double hsv[2][256][3]; /* of the copies' palettes, the first one's hue moved next to the second's */
   /* what clear_cp + flam3_add_xforms leave for the copies' shape: a result that already has
    * that shape is reset by copying this over it instead of freeing, allocating and
    * initialising its xforms again for every temporal sample */
   flam3_genome blank;
} interp_cache;

static void interp_cache_clear(interp_cache *ic) {
   if (ic->i1 >= 0) {
      clear_cp(&ic->cpi[0], 0);
      clear_cp(&ic->cpi[1], 0);
      clear_cp(&ic->blank, 0);
   }
   ic->i1 = -1;
}
/* flam3.c:774-782: the result as flam3_interpolate prepares it before the blend */
static void interp_blank(flam3_genome *g, const flam3_genome *shape) {
   clear_cp(g, 1);
   if (shape->final_xform_index >= 0) {
      flam3_add_xforms(g, shape->num_xforms - 1, 0, 0);
      flam3_add_xforms(g, 1, 0, 1);
   } else
      flam3_add_xforms(g, shape->num_xforms, 0, 0);
}

static void f3b_interpolate_cached(flam3_genome *cps, int ncps, double time, flam3_genome *result, interp_cache *ic) {
   const int have = flam3_align && flam3_interpolate_n && flam3_add_xforms && rgb2hsv && hsv2rgb;
   int i1 = 0;
   if (ncps >= 2) { /* flam3.c:729-743 */
      if (cps[0].time >= time) i1 = 0;
      else if (cps[ncps - 1].time <= time) i1 = ncps - 2;
      else {
         while (cps[i1].time < time) i1++;
         i1--;
      }
   }
   if (!have || !ic || ncps < 2 || cps[i1].interpolation != flam3_interpolation_linear || getenv("F3B_INTERP_REF")) {
      flam3_interpolate(cps, ncps, time, 0, result);
      return;
   }
   const int i2 = i1 + 1;
   if (ic->i1 != i1) {
      interp_cache_clear(ic);
      memset(ic->cpi, 0, sizeof(ic->cpi));
      flam3_align(&ic->cpi[0], &cps[i1], 2);
      ic->i1 = i1;
      ic->pal_mode = ic->cpi[0].palette_interpolation;
      if (ic->pal_mode != flam3_palette_interpolation_sweep) {
         for (int i = 0; i < 256; i++) {
            rgb2hsv(ic->cpi[0].palette[i].color, ic->hsv[0][i]);
            rgb2hsv(ic->cpi[1].palette[i].color, ic->hsv[1][i]);
            if (ic->pal_mode == flam3_palette_interpolation_hsv_circular) { /* the shorter way round */
               const double dh = ic->hsv[1][i][0] - ic->hsv[0][i][0];
               if (dh > 3.0) ic->hsv[0][i][0] += 6.0;
               else if (dh < -3.0) ic->hsv[0][i][0] -= 6.0;
            }
         }
         ic->cpi[0].palette_interpolation = flam3_palette_interpolation_sweep;
      }
      memset(&ic->blank, 0, sizeof(ic->blank));
      interp_blank(&ic->blank, &ic->cpi[0]);
   }
   double c[2];
   c[0] = (cps[i2].time - time) / (cps[i2].time - cps[i1].time);
   c[1] = 1.0 - c[0];
   const flam3_genome *bl = &ic->blank;
   if (result->xform && result->chaos && result->num_xforms == bl->num_xforms &&
       result->final_xform_index == bl->final_xform_index) {
      /* same shape as last time (the previous temporal sample): reset in place */
      flam3_xform *xf = result->xform;
      double **chaos = result->chaos;
      int motion = 0;
      for (int i = 0; i < result->num_xforms; i++) motion |= xf[i].num_motion;
      if (motion) interp_blank(result, &ic->cpi[0]); /* (never the case for an interpolated genome) */
      else {
         *result = *bl;
         result->xform = xf;
         result->chaos = chaos;
         memcpy(xf, bl->xform, (size_t)bl->num_xforms * sizeof(flam3_xform));
         const int nstd = bl->num_xforms - (bl->final_xform_index >= 0);
         for (int i = 0; i < nstd; i++) memcpy(chaos[i], bl->chaos[i], (size_t)nstd * sizeof(double));
      }
   } else
      interp_blank(result, &ic->cpi[0]);
   result->time = time;
   result->interpolation = flam3_interpolation_linear;
   result->interpolation_type = ic->cpi[0].interpolation_type;
   flam3_interpolate_n(result, 2, ic->cpi, c, 0.0);
   result->palette_interpolation = ic->pal_mode;
   if (ic->pal_mode == flam3_palette_interpolation_sweep) return;
   /* interpolation.c:377-455 with the conversions taken out of the loop */
   double rgb_fraction = 0.0;
   if (ic->pal_mode == flam3_palette_interpolation_rgb) rgb_fraction = 1.0;
   else if (ic->pal_mode == flam3_palette_interpolation_hsv_circular) rgb_fraction = ic->cpi[0].hsv_rgb_palette_blend;
   for (int i = 0; i < 256; i++) {
      const double *p0 = ic->cpi[0].palette[i].color, *p1 = ic->cpi[1].palette[i].color;
      double new_rgb[3], new_hsv[3], back[3];
      for (int k = 0; k < 3; k++) {
         new_rgb[k] = c[0] * p0[k] + c[1] * p1[k];
         new_hsv[k] = c[0] * ic->hsv[0][i][k] + c[1] * ic->hsv[1][i][k];
      }
      double count = c[0] * p0[3] + c[1] * p1[3];
      if (p0[3] == 1.0 && p1[3] == 1.0) count = 1.0;
      double index = c[0] * ic->cpi[0].palette[i].index + c[1] * ic->cpi[1].palette[i].index;
      hsv2rgb(new_hsv, back);
      double *out = result->palette[i].color;
      for (int k = 0; k < 3; k++) out[k] = rgb_fraction * new_rgb[k] + (1.0 - rgb_fraction) * back[k];
      out[3] = count;
      /* (the reference's -ffast-math build clips with max / min instructions: a -0.0 that comes
       * out of the blend leaves as +0.0) */
      for (int k = 0; k < 4; k++) {
         out[k] = out[k] > 0.0 ? out[k] : 0.0;
         out[k] = out[k] < 1.0 ? out[k] : 1.0;
      }
      index = index > 0.0 ? index : 0.0;
      result->palette[i].index = index < 255.0 ? index : 255.0;
   }
}

/* for the parity test: the genome of one time, through the cached path */
int f3b_debug_interpolate(void *cps, int ncps, double time, void *result) {
   interp_cache *ic = (interp_cache *)calloc(1, sizeof(interp_cache));
   if (!ic) return 1;
   ic->i1 = -1;
   /* twice, as a packer thread goes from one temporal sample to the next: the second call finds
    * the cache filled and the result shaped (F3B_DEBUG_INTERP_COLD=1: once, from scratch) */
   if (!getenv("F3B_DEBUG_INTERP_COLD"))
      f3b_interpolate_cached((flam3_genome *)cps, ncps, time - 0.37, (flam3_genome *)result, ic);
   f3b_interpolate_cached((flam3_genome *)cps, ncps, time, (flam3_genome *)result, ic);
   const int fast = ic->i1 >= 0;
   interp_cache_clear(ic);
   free(ic);
   return fast ? 0 : 2; /* 2: went through flam3_interpolate */
}

static void pack_one(pack_job *j, int ts, uint32_t *scratch, flam3_genome *cp, interp_cache *ic) {
   packed_sample *ps = &j->out[ts];
   const int fi = j->batch_num * j->nts + ts;
   if (!scratch) {
      ps->rc = 2;
      return;
   }
   f3b_interpolate_cached(j->spec->genomes, j->spec->ngenomes, j->spec->time + j->tdeltas[fi], cp, ic);
   if (f3b_camera_setup(cp, j->g, &ps->cam)) {
      ps->rc = 3;
      return;
   }
   ps->rc = f3b_pack_sample(cp, &ps->cam, j->tfilter[fi], j->cmap_size, j->floor_adds, scratch,
                            F3B_MAX_BLOB_WORDS, &ps->used, &ps->alpha_ok);
   if (ps->rc) return;
   if (j->slot_words && ps->used <= j->slot_words) {
      ps->words = j->slots + (size_t)ts * j->slot_words;
      ps->owned = 0;
   } else {
      ps->words = (uint32_t *)malloc(ps->used * 4);
      ps->owned = 1;
      if (!ps->words) {
         ps->rc = 2;
         return;
      }
   }
   memcpy(ps->words, scratch, ps->used * 4);
   ps->vibrancy = cp->vibrancy;
   ps->gamma = cp->gamma;
   ps->background[0] = cp->background[0];
   ps->background[1] = cp->background[1];
   ps->background[2] = cp->background[2];
   ps->contrast = cp->contrast;
   ps->brightness = cp->brightness;
   ps->estimator_curve = cp->estimator_curve;
   ps->gam_lin_thresh = cp->gam_lin_thresh;
}

static void *pack_worker(void *arg) {
   pack_job *j = (pack_job *)arg;
   uint32_t *scratch = (uint32_t *)malloc(F3B_MAX_BLOB_WORDS * 4);
   interp_cache *ic = (interp_cache *)calloc(1, sizeof(interp_cache));
   if (ic) ic->i1 = -1;
   flam3_genome cp;
   memset(&cp, 0, sizeof(cp));
   for (;;) {
      const int ts = __sync_fetch_and_add(&j->next, 1);
      if (ts >= j->nts) break;
      pack_one(j, ts, scratch, &cp, ic);
   }
   clear_cp(&cp, 0);
   if (ic) interp_cache_clear(ic);
   free(ic);
   free(scratch);
   return NULL;
}

/* The packers are threads that stay: an animation frame is 2 ms, creating and joining 15
 * threads for each took a fifth of it. One pool per process; a flam3_render that finds it busy
 * (several calls at once, F3B_CONTEXTS) packs on its own thread. */
static struct {
   pthread_mutex_t mu;
   pthread_cond_t wake, done;
   int nworkers;      /* threads created so far */
   unsigned long gen; /* job generation */
   pack_job *job;
   int want, active;  /* helpers asked for / still packing */
   int busy;          /* a job is posted and not waited for yet */
} g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, 0, 0, NULL, 0, 0, 0};

static void *pool_main(void *arg) {
   const int me = (int)(size_t)arg;
   unsigned long seen = 0;
   pthread_mutex_lock(&g_pool.mu);
   for (;;) {
      while (g_pool.gen == seen) pthread_cond_wait(&g_pool.wake, &g_pool.mu);
      seen = g_pool.gen;
      if (me >= g_pool.want) continue;
      pack_job *job = g_pool.job;
      pthread_mutex_unlock(&g_pool.mu);
      pack_worker(job);
      pthread_mutex_lock(&g_pool.mu);
      if (--g_pool.active == 0) pthread_cond_signal(&g_pool.done);
   }
   return NULL;
}
/* hands the job to up to `helpers` pool threads; returns how many took it (0: pool busy) */
static int pool_post(pack_job *job, int helpers) {
   pthread_mutex_lock(&g_pool.mu);
   if (g_pool.busy) {
      pthread_mutex_unlock(&g_pool.mu);
      return 0;
   }
   while (g_pool.nworkers < helpers) {
      pthread_t th;
      pthread_attr_t at;
      pthread_attr_init(&at);
      pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
      const int rc = pthread_create(&th, &at, pool_main, (void *)(size_t)g_pool.nworkers);
      pthread_attr_destroy(&at);
      if (rc) break;
      g_pool.nworkers++;
   }
   if (helpers > g_pool.nworkers) helpers = g_pool.nworkers;
   if (helpers > 0) {
      g_pool.job = job;
      g_pool.want = g_pool.active = helpers;
      g_pool.busy = 1;
      g_pool.gen++;
      pthread_cond_broadcast(&g_pool.wake);
   }
   pthread_mutex_unlock(&g_pool.mu);
   return helpers;
}
static void pool_wait(void) {
   pthread_mutex_lock(&g_pool.mu);
   while (g_pool.active) pthread_cond_wait(&g_pool.done, &g_pool.mu);
   g_pool.busy = 0;
   pthread_mutex_unlock(&g_pool.mu);
}

/* The reference seeds one ISAAC context per worker thread from the frame's context, 16 words
 * each, for every temporal sample (rect.c:858-865): spec->rc ends 16 * nthreads * samples
 * words further on, which is part of the call's contract. ISAAC cannot jump: a pool of 16
 * words costs ~0.1 us of strictly serial work, 1.5 ms for a frame of 1000 temporal samples
 * at nthreads = 16 - most of what an animation frame took on the host. So only the first 16
 * words of a batch become seed material for the GPU streams (one thread's worth of the
 * reference's seeding), and the rest of the advance runs in a thread of its own beside the
 * packing, the launches and the GPU work; it is joined before the next batch draws and before
 * flam3_render returns. */
typedef struct {
   pthread_t th;
   int active;
   randctx *rc;
   unsigned long n;
} rc_advance;
static __thread rc_advance t_advance;

static void *rc_advance_main(void *arg) {
   rc_advance *a = (rc_advance *)arg;
   f3b_isaac_skip(a->rc, a->n);
   return NULL;
}
static void rc_advance_join(void) {
   if (t_advance.active) pthread_join(t_advance.th, NULL);
   t_advance.active = 0;
}
static void rc_advance_start(randctx *rc, unsigned long n) {
   rc_advance_join();
   t_advance.rc = rc;
   t_advance.n = n;
   if (n >= 16384 && pthread_create(&t_advance.th, NULL, rc_advance_main, &t_advance) == 0) t_advance.active = 1;
   else f3b_isaac_skip(rc, n);
}
static uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
   z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
   z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
   return z ^ (z >> 31);
}
/* seeds of the batch's temporal samples; spec->rc is final once rc_advance_join() has run */
static void draw_sample_seeds(flam3_frame *spec, int first_sample, int nts, int nthreads, uint64_t *seeds) {
   rc_advance_join();
   uint64_t h = 0x9e3779b97f4a7c15ull;
   for (int k = 0; k < 16; k++) h = mix64(h ^ ((uint64_t)f3b_isaac_next(&spec->rc) & 0xffffffffull));
   for (int ts = 0; ts < nts; ts++) seeds[ts] = mix64(h + 0x9e3779b97f4a7c15ull * (uint64_t)(first_sample + ts + 1));
   rc_advance_start(&spec->rc, 16ul * (unsigned long)nthreads * (unsigned long)nts - 16ul);
}

static int pack_batch(flam3_frame *spec, const f3b_frame_geom *g, const double *tfilter,
                      const double *tdeltas, int batch_num, int nts, int cmap_size, int floor_adds,
                      packed_batch *out, uint64_t *seeds, int rc_threads, blob_buf *blobs) {
   packed_free(out);
   out->s = (packed_sample *)calloc((size_t)nts, sizeof(packed_sample));
   if (!out->s) return 1;
   out->n = nts;
   pack_job job = {spec, g, tfilter, tdeltas, batch_num, nts, cmap_size, floor_adds, 0, out->s, NULL, 0};
   {
      /* Sample 0 sizes the slots of the batch buffer (the samples of a frame are the same
       * genome at different times: same size, give or take padding); the others are then
       * packed straight into the page-locked buffer by the workers. */
      uint32_t *scratch = (uint32_t *)malloc(F3B_MAX_BLOB_WORDS * 4);
      flam3_genome cp;
      memset(&cp, 0, sizeof(cp));
      pack_one(&job, 0, scratch, &cp, NULL);
      clear_cp(&cp, 0);
      free(scratch);
      job.next = 1;
      if (out->s[0].rc == 0) {
         const size_t slot = ((out->s[0].used + 3) & ~(size_t)3) + 64;
         blobs->used = 0;
         if (blob_reserve(blobs, slot * (size_t)nts)) return 1;
         memcpy(blobs->words, out->s[0].words, out->s[0].used * 4);
         free(out->s[0].words);
         out->s[0].words = blobs->words;
         out->s[0].owned = 0;
         job.slots = blobs->words;
         job.slot_words = slot;
         out->slot_words = slot;
         blobs->used = slot * (size_t)nts;
      }
   }
   int nthreads = 1;
   if (nts >= 32) {
      const char *e = getenv("F3B_HOST_THREADS");
      long nproc = sysconf(_SC_NPROCESSORS_ONLN);
      nthreads = e ? atoi(e) : (int)(nproc > 16 ? 16 : nproc);
      if (nthreads > nts / 8) nthreads = nts / 8;
      if (nthreads < 1) nthreads = 1;
   }
   if (nthreads > 64) nthreads = 64;
   const int helpers = nthreads > 1 ? pool_post(&job, nthreads - 1) : 0;
   draw_sample_seeds(spec, batch_num * nts, nts, rc_threads, seeds);
   pack_worker(&job);
   if (helpers) pool_wait();
   return 0;
}

static int render_locked(f3b_ctx *ctx, flam3_frame *spec, void *out, int field, int nchan,
                         int transp, stat_struct *stats, int force_count_plane) {
   const time_t tstart = time(NULL);
   g_progress_t0 = now_ms();
   const int timing = getenv("F3B_TIMING") != NULL;
   double t_setup = now_ms(), t_pack = 0, t_serial = 0, t_iter = 0, t_dens = 0, t_fin = 0, t0;
   int rv = 1;
   enum { MODE_SINGLE, MODE_PEER_LOCAL, MODE_PEER_IPC, MODE_CALLBACKS } mode = MODE_SINGLE;
   int NR = 1, nlocal = 1, rank0 = 0; /* ranks sharing the frame, contexts of this process, first rank here */
   f3b_ctx **lc = g_group.ctx;
   int verbose = spec->verbose;
   const int nthreads = spec->nthreads > 0 ? spec->nthreads : 1;
   const int bytes_per_channel = spec->bytes_per_channel;
   const char *ai = getenv("apo_palette");
   const int cmap_size = 256 - (ai ? atoi(ai) : 0);

   double *filter = NULL, *temporal_filter = NULL, *temporal_deltas = NULL;
   blob_buf blobs = g_blobs; /* handed back to g_blobs at the end */
   f3b_job *jobs = NULL;
   uint64_t *sample_seeds = NULL;
   packed_batch packed = {NULL, 0, 0};
   packed_sample last;
   memset(&last, 0, sizeof(last));
   f3b_de_filters de;
   int de_alive = 0;
   double last_dmap[256][3];
   flam3_genome cp;
   memset(&cp, 0, sizeof(cp));
   memset(last_dmap, 0, sizeof(last_dmap));

   stats->num_iters = 0;
   stats->badvals = 0;

   flam3_interpolate(spec->genomes, spec->ngenomes, spec->time, 0, &cp);
   const int oversample = cp.spatial_oversample;
   const double highpow = cp.highlight_power;
   const int nbatches = cp.nbatches;
   const int ntemporal_samples = cp.ntemporal_samples;

   if (nbatches < 1) {
      fprintf(stderr, "nbatches must be positive, not %d.\n", nbatches);
      goto done;
   }
   if (oversample < 1) {
      fprintf(stderr, "oversample must be positive, not %d.\n", oversample);
      goto done;
   }
   if (ntemporal_samples < 1) {
      fprintf(stderr, "temporal_samples must be positive, not %d.\n", ntemporal_samples);
      goto done;
   }

   /* output geometry, adjusted for scanline fields (rect.c:611-624) */
   f3b_frame_geom g;
   memset(&g, 0, sizeof(g));
   g.field = field;
   g.pixel_aspect_ratio = spec->pixel_aspect_ratio;
   g.oversample = oversample;
   g.image_width = cp.width;
   g.out_width = g.image_width;
   if (field) {
      g.image_height = cp.height / 2;
      if (field == flam3_field_odd)
         out = (unsigned char *)out + (size_t)nchan * bytes_per_channel * g.out_width;
      g.out_width *= 2;
   } else
      g.image_height = cp.height;

   /* spatial filter from genomes[0], NOT the interpolated cp (filters.c:219-222) */
   g.filter_width = f3b_spatial_filter_create(
       spec->genomes[0].spatial_filter_select, spec->genomes[0].spatial_oversample,
       spec->genomes[0].spatial_filter_radius, spec->pixel_aspect_ratio, field, &filter);
   if (g.filter_width < 0) {
      fprintf(stderr, "Spatial filter value is too small: %g.  Terminating.\n",
              spec->genomes[0].spatial_filter_radius);
      fprintf(stderr, "flam3_create_spatial_filter returned error: aborting\n");
      goto done;
   }

   const double sumfilt = f3b_temporal_filter_create(
       nbatches * ntemporal_samples, cp.temporal_filter_type, cp.temporal_filter_exp,
       cp.temporal_filter_width, &temporal_filter, &temporal_deltas);

   /* gutter: room for the spatial filter and for the widest DE kernel (rect.c:656-681) */
   g.gutter_width = (g.filter_width - oversample) / 2;
   int max_de = 0;
   for (int i = 0; i < spec->ngenomes; i++) {
      const int w = (int)ceil(spec->genomes[i].estimator) * oversample;
      if (w > max_de) max_de = w;
   }
   if (max_de > 0) max_de += oversample - 1;
   if (max_de > g.gutter_width) {
      g.de_offset = max_de - g.gutter_width;
      g.gutter_width = max_de;
   } else
      g.de_offset = 0;
   g.height = oversample * g.image_height + 2 * g.gutter_width;
   g.width = oversample * g.image_width + 2 * g.gutter_width;

   /* ---- how many GPUs share this frame, and how ----
    * PEER_LOCAL: the contexts of this process (F3B_DEVICES), rank i = context i.
    * PEER_IPC:   one context here, the other ranks in other processes (f3b_render_set_peers).
    * legacy:     one context, host callbacks sum the histograms and gather the strips
    *             (f3b_render_set_exchange / f3b_render_set_strip_gather). */
   if (g_group.n > 1 && !g_inject) {
      mode = MODE_PEER_LOCAL;
      NR = nlocal = g_group.n;
   } else if (g_npeers > 1 && g_nranks == g_npeers && g_barrier && !g_inject) {
      mode = MODE_PEER_IPC;
      NR = g_nranks;
      rank0 = g_rank;
   } else if (g_nranks > 1) {
      mode = MODE_CALLBACKS;
      NR = g_nranks;
      rank0 = g_rank;
   }
   const int peer_mode = (mode == MODE_PEER_LOCAL || mode == MODE_PEER_IPC);
   /* row strips of the post-pass: rank r handles image rows [r*per, (r+1)*per) */
   const int strips = peer_mode || (mode == MODE_CALLBACKS && g_gather != NULL && g_exchange != NULL);
   const int strip_rows_per = (g.image_height + NR - 1) / NR;
   int strip_row0 = rank0 * strip_rows_per, strip_row1 = strip_row0 + strip_rows_per;
   if (strip_row0 > g.image_height) strip_row0 = g.image_height;
   if (strip_row1 > g.image_height) strip_row1 = g.image_height;

   /* Can the alpha plane double as the hit-count plane? (f3b_pack.c, dmap section) */
   int count_plane = force_count_plane;
   for (int i = 0; i < nbatches * ntemporal_samples && !count_plane; i++)
      if (temporal_filter[i] != 1.0) count_plane = 1;
   for (int i = 0; i < spec->ngenomes && !count_plane; i++) {
      if (spec->genomes[i].palette_mode == flam3_palette_mode_linear) count_plane = 1;
      for (int j = 0; j < 256 && !count_plane; j++)
         if (spec->genomes[i].palette[j].color[3] != 1.0) count_plane = 1;
   }

   for (int i = 0; i < nlocal; i++)
      if (f3b_frame_begin(lc[i], g.width, g.height, count_plane, spec->bits)) {
         fprintf(stderr, "render_rectangle (B200): cannot allocate %dx%d buckets: %s\n", g.width,
                 g.height, f3b_last_error(lc[i]));
         goto done;
      }
   if (mode == MODE_PEER_IPC) {
      /* the peers were told where this rank's buckets are: they must still be there */
      if (!g_peer_own_color) g_peer_own_color = f3b_hist_color_ptr(ctx);
      if (g_peer_own_color != f3b_hist_color_ptr(ctx)) {
         fprintf(stderr, "flam3_render (B200): the bucket planes moved (the frame grew) after their "
                         "address was given to the other ranks; exchange the addresses again.\n");
         goto done;
      }
   }

   if (verbose) fprintf(stderr, "chaos: ");
   const time_t progress_began = time(NULL);

   double vibrancy = 0.0, gamma = 0.0, background[3] = {0.0, 0.0, 0.0};
   int vib_gam_n = 0;
   f3b_camera cam;
   memset(&cam, 0, sizeof(cam));
   jobs = (f3b_job *)calloc((size_t)ntemporal_samples, sizeof(f3b_job));
   sample_seeds = (uint64_t *)calloc((size_t)ntemporal_samples, sizeof(uint64_t));
   if (!jobs || !sample_seeds) goto done;

   for (int batch_num = 0; batch_num < nbatches; batch_num++) {
      const double de_time = spec->time + temporal_deltas[batch_num * ntemporal_samples];

      if (batch_num > 0)
         for (int i = 0; i < nlocal; i++)
            if (f3b_hist_clear(lc[i])) {
               ctx = lc[i];
               goto gpu_error;
            }

      /* DE parameters come from the genome at the batch's first time (rect.c:726-751) */
      flam3_interpolate(spec->genomes, spec->ngenomes, de_time, 0, &cp);
      if (cp.estimator < 0.0 || cp.estimator_minimum < 0.0) {
         fprintf(stderr, "density estimator filter widths must be >= 0\n");
         goto done;
      }
      if (spec->bits <= 32 && cp.estimator > 0.0) {
         fprintf(stderr, "warning: density estimation disabled with %d bit buffers.\n", spec->bits);
         cp.estimator = 0.0;
      }
      memset(&de, 0, sizeof(de));
      de_alive = 0;
      if (cp.estimator > 0.0) {
         if (de_filters_cached(cp.estimator, cp.estimator_minimum, cp.estimator_curve, oversample,
                               &de)) {
            fprintf(stderr, "de.kernel_size returned 0 - aborting.\n");
            goto done;
         }
         de_alive = 1;
      }

      /* ---- pack every temporal sample of this batch (rect.c:754-884). Interpolating and
       * packing a sample is independent host work (~45 us), so with many temporal samples
       * (flam3-animate: 1000 per frame) it is spread over a few host threads. ---- */
      for (int i = 0; i < nlocal; i++) /* the previous upload has left the buffer */
         if (f3b_inputs_consumed(lc[i])) {
            ctx = lc[i];
            goto gpu_error;
         }
      blobs.used = 0;
      const uint32_t max_chain = spec->sub_batch_size > 0 ? (uint32_t)spec->sub_batch_size : 10000u;
      uint64_t seed = 0x243f6a8885a308d3ull;
      if (batch_num == 0) t_setup = now_ms() - t_setup;
      t0 = now_ms();
      if (pack_batch(spec, &g, temporal_filter, temporal_deltas, batch_num, ntemporal_samples,
                     cmap_size, spec->bits != 64, &packed, sample_seeds, nthreads, &blobs)) {
         fprintf(stderr, "flam3_render (B200): out of memory packing temporal samples\n");
         goto done;
      }
      t_pack += now_ms() - t0;
      t0 = now_ms();
      for (int ts = 0; ts < ntemporal_samples; ts++) {
         const packed_sample *ps = &packed.s[ts];
         const size_t used = ps->used;
         if (ps->rc == 3) goto done; /* camera error, message already printed */
         if (ps->rc == 2) {
            fprintf(stderr, "flam3_render (B200): genome too large for the packed format.\n");
            goto done;
         }
         if (ps->rc) {
            fprintf(stderr, "create xform distrib returned error: aborting.\n");
            goto done;
         }
         if (!ps->alpha_ok && !count_plane) {
            rv = -2; /* caller retries with a count plane */
            goto done;
         }
         size_t blob_word_offset;
         if (ps->owned) { /* did not fit its slot: appended behind the slots */
            if (blob_reserve(&blobs, used + 4)) goto done;
            memcpy(blobs.words + blobs.used, ps->words, used * 4);
            blob_word_offset = blobs.used;
            blobs.used += (used + 3) & ~(size_t)3;
         } else
            blob_word_offset = (size_t)ts * packed.slot_words;
         cam = ps->cam;

         /* sample budget (rect.c:830-835) */
         const double nsamples = cam.sample_density * g.image_width * g.image_height;
         const double batch_size = nsamples / (nbatches * ntemporal_samples);
         stats->num_iters += batch_size;

         const uint64_t h = sample_seeds[ts]; /* drawn from spec->rc in pack_batch */
         if (ts == 0) seed ^= h;

         f3b_job *jb = &jobs[ts];
         jb->blob_offset = (uint64_t)blob_word_offset * 4;
         jb->n_iters = batch_size > 0 ? (uint64_t)batch_size : 0;
         jb->n_chains = 0; /* planned by f3b_iterate, which knows the kernel class of the batch */
         jb->fuse = spec->earlyclip ? 100 : 15; /* FUSE_28 : FUSE_27, rect.c:39-40,266 */
         jb->first_group = 0;
         jb->sample_id = (uint32_t)(h ^ (h >> 32));

         vibrancy += ps->vibrancy;
         gamma += ps->gamma;
         background[0] += ps->background[0];
         background[1] += ps->background[1];
         background[2] += ps->background[2];
         vib_gam_n++;
         if (ts == ntemporal_samples - 1) {
            /* the reference's k1/k2, DE curve and gamma threshold come from the cp of the
             * batch's last temporal sample (rect.c:933-937, 997, 1076) */
            last = *ps;
            if (batch_num == nbatches - 1) {
               /* (through the batch buffer: an append above may have moved it) */
               const uint32_t *pw = blobs.words + blob_word_offset;
               const float *pal = (const float *)(pw + pw[F3B_H_OFF_PALETTE]);
               for (int j = 0; j < 256; j++)
                  for (int k = 0; k < 3; k++) last_dmap[j][k] = pal[j * 4 + k];
            }
         }
      }

      if (poll_progress(spec, 100.0 * batch_num / nbatches, 0)) {
         if (verbose) fprintf(stderr, "\naborted!\n");
         rv = 0;
         goto done;
      }

      t_serial += now_ms() - t0;
      t0 = now_ms();
      /* ---- chaos game: one launch for the whole batch ---- */
      if (g_inject) {
         if (g_inject_n != (size_t)g.width * g.height) {
            fprintf(stderr, "flam3_render (B200): injected histogram has %zu buckets, frame needs "
                            "%dx%d\n", g_inject_n, g.width, g.height);
            goto done;
         }
         if (f3b_hist_upload_u32(ctx, g_inject)) goto gpu_error;
      }
      if (g_inject && !g_inject_preload) {
         /* post-pass only */
      } else {
         /* With a progress callback installed (GUIs; the CLIs have none) a long batch is cut
          * into slices of >= 2.5e8 iterations per GPU - slice s of S is "rank r + s N of S N"
          * of the same chain-group sharding, so the image does not change - and the callback
          * is polled between them (percent, ETA, abort, pause: rect.c:324-373). Without a
          * callback the batch is one launch sequence and nothing waits on the host. */
         int nslices = 1;
         if (spec->progress) {
            double total = 0.0;
            for (int ts = 0; ts < ntemporal_samples; ts++) total += (double)jobs[ts].n_iters;
            nslices = (int)(total / (2.5e8 * NR));
            if (nslices < 1) nslices = 1;
            if (nslices > 64) nslices = 64;
         }
         for (int sl = 0; sl < nslices; sl++) {
            for (int i = 0; i < nlocal; i++) {
               /* the plan is the same for every call: only first_group is filled in */
               for (int ts = 0; ts < ntemporal_samples; ts++) jobs[ts].first_group = 0;
               if (f3b_iterate(lc[i], blobs.words, blobs.used * 4, jobs, ntemporal_samples, seed,
                               spec->bits != 64, max_chain, (uint32_t)nthreads, rank0 + i + sl * NR,
                               NR * nslices)) {
                  ctx = lc[i];
                  goto gpu_error;
               }
            }
            if (nslices > 1) {
               for (int i = 0; i < nlocal; i++)
                  if (f3b_sync(lc[i])) {
                     ctx = lc[i];
                     goto gpu_error;
                  }
               if (poll_progress(spec, 100.0 * (batch_num + (sl + 1.0) / nslices) / nbatches, 0)) {
                  if (verbose) fprintf(stderr, "\naborted!\n");
                  rv = 0;
                  goto done;
               }
            }
         }
      }
      if (mode == MODE_CALLBACKS && g_exchange) {
         if (!g_stream_ordered && f3b_sync(ctx)) goto gpu_error;
         (*g_exchange)(ctx, g_exchange_user);
      }

      t_iter += now_ms() - t0;
      t0 = now_ms();
      /* ---- log scale + density estimation (rect.c:933-1058); cp/cam are those of the
       * batch's LAST temporal sample, as in the reference ---- */
      f3b_density_desc dd;
      memset(&dd, 0, sizeof(dd));
      dd.k1 = (last.contrast * last.brightness * 255.0 * 268.0 * (1.0 / nbatches)) / 256;
      const double area = g.image_width * g.image_height / (cam.ppux * cam.ppuy);
      dd.k2 = (oversample * oversample * nbatches) /
              (last.contrast * area * 255.0 * cam.sample_density * sumfilt);
      dd.oversample = oversample;
      dd.int_accum = (spec->bits == 32);
      if (de_alive && de.max_filter_index > 0) {
         dd.n_kernels = de.n_kernels;
         dd.kernel_size = de.kernel_size;
         dd.max_filtered_counts = de.max_filtered_counts;
         dd.max_filter_index = de.max_filter_index;
         dd.curve = last.estimator_curve;
         dd.coefs = de.coefs;
         dd.widths = de.widths;
      }
      if (poll_progress(spec, 0.0, 1)) {
         if (verbose) fprintf(stderr, "\naborted!\n");
         rv = 0;
         goto done;
      }
      if (peer_mode) {
         /* Reduce-scatter by image strip, fused: when every rank's chaos game is through,
          * rank r sums all ranks' buckets of the rows its strip needs - the accumulator rows
          * under its pixels' filter footprints plus the reach of the widest DE kernel -
          * straight out of the peers' memory, and nobody clears buckets before all have read. */
         if (mode == MODE_PEER_LOCAL ? f3b_group_barrier(lc, nlocal) : ((*g_barrier)(ctx, g_barrier_user), 0))
            goto gpu_error;
         f3b_peer_hist peers[F3B_MAX_LOCAL];
         for (int r = 0; r < NR; r++) {
            if (mode == MODE_PEER_LOCAL) {
               peers[r].color = f3b_hist_color_ptr(lc[r]);
               peers[r].count = f3b_hist_count_ptr(lc[r]);
            } else
               peers[r] = g_peers[r];
         }
         const int halo = max_de + oversample;
         for (int i = 0; i < nlocal; i++) {
            int s0 = (rank0 + i) * strip_rows_per, s1 = s0 + strip_rows_per;
            if (s0 > g.image_height) s0 = g.image_height;
            if (s1 > g.image_height) s1 = g.image_height;
            if (s1 <= s0) continue;
            const int a0 = g.de_offset + s0 * oversample;
            const int a1 = g.de_offset + (s1 - 1) * oversample + g.filter_width;
            if (f3b_hist_reduce_rows(lc[i], peers, NR, a0 - halo, a1 + halo)) {
               ctx = lc[i];
               goto gpu_error;
            }
         }
         if (mode == MODE_PEER_LOCAL ? f3b_group_barrier(lc, nlocal) : ((*g_barrier)(ctx, g_barrier_user), 0))
            goto gpu_error;
         for (int i = 0; i < nlocal; i++) {
            int s0 = (rank0 + i) * strip_rows_per, s1 = s0 + strip_rows_per;
            if (s0 > g.image_height) s0 = g.image_height;
            if (s1 > g.image_height) s1 = g.image_height;
            if (s1 <= s0) continue;
            const int a0 = g.de_offset + s0 * oversample;
            const int a1 = g.de_offset + (s1 - 1) * oversample + g.filter_width;
            if (f3b_density_rows(lc[i], &dd, a0, a1)) {
               ctx = lc[i];
               goto gpu_error;
            }
         }
      } else if (strips) {
         /* accumulator rows under the filter footprints of this rank's image rows */
         const int a0 = g.de_offset + strip_row0 * oversample;
         const int a1 = g.de_offset + (strip_row1 - 1) * oversample + g.filter_width;
         if (strip_row1 > strip_row0 && f3b_density_rows(ctx, &dd, a0, a1)) goto gpu_error;
      } else if (f3b_density(ctx, &dd))
         goto gpu_error;
      de_alive = 0; /* the bank belongs to the cache */
      t_dens += now_ms() - t0;
   }
   t0 = now_ms();

   if (verbose) {
      fprintf(stderr, "\rchaos: 100.0%%  took: %ld seconds   \n", (long)(time(NULL) - progress_began));
      fprintf(stderr, "filtering...");
   }

   /* ---- spatial filter + tone map (rect.c:1067-1247) ---- */
   {
      f3b_final_desc fd;
      memset(&fd, 0, sizeof(fd));
      fd.image_width = g.image_width;
      fd.image_height = g.image_height;
      fd.oversample = oversample;
      fd.de_offset = g.de_offset;
      fd.filter_width = g.filter_width;
      fd.filter = filter;
      fd.earlyclip = spec->earlyclip;
      fd.nchan = nchan;
      fd.transp = transp;
      fd.bytes_per_channel = bytes_per_channel;
      fd.gamma = gamma / vib_gam_n;
      fd.vibrancy = vibrancy / vib_gam_n;
      for (int k = 0; k < 3; k++) fd.background[k] = background[k] / (vib_gam_n / 256.0);
      fd.highlight_power = highpow;
      fd.gam_lin_thresh = last.gam_lin_thresh;
      const size_t stride = (size_t)g.out_width * nchan * bytes_per_channel;
      if (peer_mode) {
         /* every rank's image rows go straight into the image of the rank that owns `out` */
         void *image = g_peer_image0;
         const int owner_here = (rank0 == 0);
         if (owner_here && f3b_image_reserve(lc[0], &fd, &image)) goto gpu_error;
         if (mode == MODE_PEER_IPC && owner_here && g_peer_image0 && image != g_peer_image0) {
            fprintf(stderr, "flam3_render (B200): the image buffer moved after its address was given "
                            "to the other ranks; exchange the addresses again.\n");
            goto done;
         }
         for (int i = 0; i < nlocal; i++) {
            int s0 = (rank0 + i) * strip_rows_per, s1 = s0 + strip_rows_per;
            if (s0 > g.image_height) s0 = g.image_height;
            if (s1 > g.image_height) s1 = g.image_height;
            if (s1 <= s0) continue;
            if (f3b_finish_rows_into(lc[i], &fd, s0, s1, image)) {
               ctx = lc[i];
               goto gpu_error;
            }
         }
         if (mode == MODE_PEER_LOCAL ? f3b_group_barrier(lc, nlocal) : ((*g_barrier)(ctx, g_barrier_user), 0))
            goto gpu_error;
         if (owner_here && !g_device_only && f3b_image_download(lc[0], &fd, out, stride)) goto gpu_error;
      } else if (strips) {
         void *image = NULL;
         if (f3b_finish_device_rows(ctx, &fd, strip_row0, strip_row1, strip_rows_per * g_nranks, &image) ||
             (!g_stream_ordered && f3b_sync(ctx)))
            goto gpu_error;
         (*g_gather)(ctx, image, (size_t)g.image_width * nchan * bytes_per_channel, strip_rows_per,
                     g_gather_user);
         if (!g_device_only && f3b_image_download(ctx, &fd, out, stride)) goto gpu_error;
      } else if (g_device_only) {
         if (f3b_finish_device(ctx, &fd, NULL)) goto gpu_error;
      } else if (f3b_finish(ctx, &fd, out, stride))
         goto gpu_error;
   }

   for (int i = 0; i < nlocal; i++) {
      double bv = 0.0;
      if (f3b_badvals(lc[i], &bv)) {
         ctx = lc[i];
         goto gpu_error;
      }
      stats->badvals += bv;
   }

   if (getenv("insert_palette") && out && !g_device_only) { /* debug aid of the reference, rect.c:1267-1279 */
      int ph = 100;
      if (ph >= g.image_height) ph = g.image_height;
      for (int j = 0; j < ph; j++)
         for (int i = 0; i < g.image_width; i++) {
            unsigned char *p = (unsigned char *)out + (size_t)nchan * (i + (size_t)j * g.out_width);
            for (int k = 0; k < 3; k++) p[k] = (unsigned char)last_dmap[i * 256 / g.image_width][k];
         }
   }
   t_fin = now_ms() - t0;
   if (timing)
      fprintf(stderr, "f3b timing (host ms): setup %.2f pack %.2f jobs %.2f iterate-submit %.2f "
                      "density-submit %.2f finish+download %.2f\n", t_setup, t_pack, t_serial, t_iter,
              t_dens, t_fin);
   rv = 0;
   goto done;

gpu_error:
   fprintf(stderr, "flam3_render (B200): %s\n", f3b_last_error(ctx));
   rv = 1;

done:
   free(temporal_filter);
   free(temporal_deltas);
   free(filter);
   for (int i = 0; i < nlocal; i++)
      if (f3b_inputs_consumed(lc[i])) rv = 1;
   blobs.used = 0;
   g_blobs = blobs;
   free(jobs);
   free(sample_seeds);
   packed_free(&packed);
   clear_cp(&cp, 0);
   stats->render_seconds = (int)(time(NULL) - tstart);
   return rv;
}

int flam3_render(flam3_frame *spec, void *out, int field, int nchan, int transp,
                 stat_struct *stats) {
   if (spec->bits != 32 && spec->bits != 33 && spec->bits != 64) {
      fprintf(stderr, "flam3: bits must be 32, 33, or 64 not %d.\n", spec->bits);
      return 1;
   }
   static __thread double t_last_return = 0.0;
   const double t_enter = now_ms();
   f3b_ctx *ctx = slot_acquire(); /* waits while all F3B_CONTEXTS slots are rendering */
   int rv = 1;
   if (ctx) {
      const randctx rc0 = spec->rc;
      rv = render_locked(ctx, spec, out, field, nchan, transp, stats, g_inject != NULL);
      rc_advance_join(); /* spec->rc is where the reference would have left it */
      if (rv == -2) { /* a palette with alpha appeared mid-frame: redo with a count plane */
         spec->rc = rc0;
         rv = render_locked(ctx, spec, out, field, nchan, transp, stats, 1);
         rc_advance_join();
      }
      if (rv < 0) rv = 1;
      slot_release();
   }
   if (getenv("F3B_TIMING"))
      fprintf(stderr, "f3b timing: flam3_render %.2f ms, %.2f ms in the caller since the previous call\n",
              now_ms() - t_enter, t_last_return > 0.0 ? t_enter - t_last_return : 0.0);
   t_last_return = now_ms();
   return rv;
}
