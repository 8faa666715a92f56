/*
 * f3b_api.cu — the C-ABI of libflam3_b200.so (include/flam3_b200.h): context, device
 * buffers, and the launches that replace render_rectangle's thread machinery
 * (rect.c:685-1247). Host-side glue only; the kernels are in f3b_iterate.cu / f3b_post.cu.
 */
#include <cuda_runtime.h>
#include <time.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "f3b_internal.h"

struct DevBuf {
   void *p = nullptr;
   size_t cap = 0;
};

#define F3B_SIDE_STREAMS 3
/* device counters (unsigned long long each): [0] bad values, [1] fullest bucket's count,
 * [2] lanes adding / [3] lanes sharing a bucket in the aggregation probe, [4] the probe's own
 * bad values (discarded), [5] spare, [6] staged values that reached 2^24 (exactness lost),
 * [7] entries of the saturation list, [8..9] the sweepers' state (three uint32: cold ticket,
 * hot entries, hot ticket; cleared before every launch) */
#define F3B_NCOUNTERS 16
#define F3B_CTR_LOST 6
#define F3B_CTR_TICKET 8

struct f3b_ctx {
   int device = 0;
   int sms = 0;
   int iter_blocks_per_sm = 0;         /* resident CTAs per SM of the SIMPLE kernels ... */
   int iter_blocks_per_sm_generic = 0; /* ... and of the generic ones */
   cudaStream_t stream = nullptr;
   char err[512] = {0};

   int width = 0, height = 0;
   F3HistLayout lay = {0, 0, 0}; /* histogram addressing */
   size_t hist_elems = 0;        /* buckets in a histogram plane */
   int count_channel = 1;
   DevBuf hist_color, hist_count, accum, image; /* hist_*: the float STAGING copies */
   DevBuf exact_color, exact_count, sat_list;   /* the buckets proper (f3b_internal.h) */
   DevBuf xfer;                 /* staging of f3b_hist_upload/download_u32 (kept across calls) */
   DevBuf hot_list, hot_bitmap; /* the sweepers' list of fast-growing staging entries */
   DevBuf fixup;                /* final pass: [0] count, [4..] pixels to redo in double */
   /* binned scatter (f3b_iterate.cu): the staging planes are ONE L2-sized tile, hits travel
    * as records through per-region queues */
   bool binned = false;
   int bin_regions = 0, bin_region_rows = 0;
   uint32_t bin_magic = 0;
   size_t bin_cap = 0, bin_tile_elems = 0;
   DevBuf bin_records, bin_pos, bin_table;
   bool bin_forced = false;     /* F3B_BINNED=1: always, else only when f3b_iterate finds it safe */
   bool bin_batch_open = false;
   uint64_t bin_batch_iters = 0;
   unsigned bin_launch_seq = 0;
   /* several GPUs on one frame: the strip of rows this context owns, summed over all ranks */
   DevBuf reduced_color, reduced_count;
   bool view_reduced = false;
   int view_row0 = 0, view_rows = 0;
   cudaEvent_t bar_ev = nullptr; /* f3b_group_barrier */
   int f64 = 0;                 /* bits 64: double buckets */
   bool staging_dirty = true;   /* staging copies may hold something (cleared by the merge) */
   bool swept = false;          /* sweepers have run since the last hist_clear */
   uint64_t lost_exact = 0, saturated = 0; /* as of the last f3b_badvals */
   DevBuf blobs, jobs, groups, de_coefs, de_radius, de_kparam, de_d2max, filter, de_kidx, de_tiles;
   std::vector<float> de_coefs_host;    /* what de_coefs / de_radius / filter hold on the device */
   std::vector<uint8_t> de_radius_host;
   std::vector<double> filter_host;
   bool accum_zero_pending = false; /* accumulators are logically zero, memset not yet issued */
   unsigned long long *d_badvals = nullptr;
   cudaStream_t side[F3B_SIDE_STREAMS] = {nullptr, nullptr, nullptr};
   /* fullest bucket's share of the frame's count plane, measured by the density pass of the
    * last frame, and which genome it belongs to (f3b_iterate decides about aggregation) */
   double hot_share = -1.0, hot_total = 0.0;
   bool hot_measured = false;   /* a pass over the whole histogram has written the fullest bucket */
   bool hot_from_merge = false; /* ... the merge of the histogram copies (this rank's share) */
   uint64_t hot_iters = 0;
   uint64_t hot_share_signature = 0, hot_signature = 0;
   cudaEvent_t inputs_ev = nullptr; /* the last f3b_iterate's uploads have left the host buffers */
   bool inputs_pending = false;
   cudaEvent_t fork_ev = nullptr, join_ev[F3B_SIDE_STREAMS] = {nullptr, nullptr, nullptr};

   /* instrumentation */
   std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
   struct Rec {
      int stage;
      int ev;
   };
   std::vector<Rec> recs;
   size_t ev_used = 0;
   double stage_ms[F3B_NSTAGES] = {0};
   uint64_t stage_launches[F3B_NSTAGES] = {0};
   uint64_t h2d_bytes = 0, d2h_bytes = 0, peer_bytes = 0;
   cudaEvent_t timer0 = nullptr, timer1 = nullptr;
};

static int fail(f3b_ctx *c, const char *fmt, ...) {
   va_list ap;
   va_start(ap, fmt);
   if (c) vsnprintf(c->err, sizeof(c->err), fmt, ap);
   va_end(ap);
   if (c && getenv("F3B_VERBOSE")) fprintf(stderr, "flam3_b200: %s\n", c->err);
   return 1;
}

#define CK(call)                                                                         \
   do {                                                                                  \
      cudaError_t e_ = (call);                                                           \
      if (e_ != cudaSuccess)                                                             \
         return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                          \
   } while (0)

static int ensure(f3b_ctx *ctx, DevBuf &b, size_t bytes) {
   if (bytes <= b.cap) return 0;
   if (b.p) cudaFree(b.p);
   b.p = nullptr;
   b.cap = 0;
   size_t want = bytes + bytes / 8 + 256;
   CK(cudaMalloc(&b.p, want));
   b.cap = want;
   return 0;
}

static int stats_collect(f3b_ctx *ctx);
static int stage_begin(f3b_ctx *ctx, int stage) {
   if (ctx->ev_used >= 1024 && stats_collect(ctx)) return 1; /* bound the event pool */
   if (ctx->ev_used == ctx->ev_pool.size()) {
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&b));
      ctx->ev_pool.push_back({a, b});
   }
   ctx->recs.push_back({stage, (int)ctx->ev_used});
   CK(cudaEventRecord(ctx->ev_pool[ctx->ev_used].first, ctx->stream));
   return 0;
}
static int stage_end(f3b_ctx *ctx, int launches) {
   const f3b_ctx::Rec &r = ctx->recs.back();
   CK(cudaEventRecord(ctx->ev_pool[ctx->ev_used].second, ctx->stream));
   ctx->ev_used++;
   ctx->stage_launches[r.stage] += (uint64_t)launches;
   return 0;
}
static int stats_collect(f3b_ctx *ctx) {
   CK(cudaStreamSynchronize(ctx->stream));
   for (const f3b_ctx::Rec &r : ctx->recs) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ctx->ev_pool[r.ev].first, ctx->ev_pool[r.ev].second));
      ctx->stage_ms[r.stage] += ms;
   }
   ctx->recs.clear();
   ctx->ev_used = 0;
   return 0;
}

extern "C" {

int f3b_abi_version(void) { return F3B_ABI_VERSION; }

int f3b_ctx_create(int device, f3b_ctx **out) {
   if (!out) return 1;
   *out = nullptr;
   f3b_ctx *ctx = new f3b_ctx();
   ctx->device = device;
   cudaError_t e = cudaSetDevice(device);
   if (e != cudaSuccess) {
      fprintf(stderr, "flam3_b200: cudaSetDevice(%d) failed: %s\n", device, cudaGetErrorString(e));
      delete ctx;
      return 1;
   }
   cudaDeviceProp prop;
   e = cudaGetDeviceProperties(&prop, device);
   if (e != cudaSuccess || prop.major < 10) {
      fprintf(stderr, "flam3_b200: device %d is not an sm_100-class GPU (%s)\n", device,
              e == cudaSuccess ? prop.name : cudaGetErrorString(e));
      delete ctx;
      return 1;
   }
   ctx->sms = prop.multiProcessorCount;
   if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
       cudaMalloc(&ctx->d_badvals, F3B_NCOUNTERS * sizeof(unsigned long long)) != cudaSuccess) {
      fprintf(stderr, "flam3_b200: context setup failed: %s\n",
              cudaGetErrorString(cudaGetLastError()));
      delete ctx;
      return 1;
   }
   cudaMemsetAsync(ctx->d_badvals, 0, F3B_NCOUNTERS * sizeof(unsigned long long), ctx->stream);
   ctx->iter_blocks_per_sm = f3b_iterate_max_blocks_per_sm(16 * 1024, true);
   if (ctx->iter_blocks_per_sm <= 0) ctx->iter_blocks_per_sm = 1;
   ctx->iter_blocks_per_sm_generic = f3b_iterate_max_blocks_per_sm(16 * 1024, false);
   if (ctx->iter_blocks_per_sm_generic <= 0) ctx->iter_blocks_per_sm_generic = 1;
   *out = ctx;
   return 0;
}

void f3b_ctx_destroy(f3b_ctx *ctx) {
   if (!ctx) return;
   cudaSetDevice(ctx->device);
   cudaStreamSynchronize(ctx->stream);
   DevBuf *bufs[] = {&ctx->exact_color, &ctx->exact_count, &ctx->sat_list, &ctx->xfer,
                     &ctx->reduced_color, &ctx->reduced_count, &ctx->bin_records, &ctx->bin_pos, &ctx->bin_table,
                     &ctx->hot_list, &ctx->hot_bitmap, &ctx->fixup,
                     &ctx->hist_color, &ctx->hist_count, &ctx->accum,    &ctx->image,
                     &ctx->blobs,      &ctx->jobs,       &ctx->groups,   &ctx->de_coefs,
                     &ctx->de_radius,  &ctx->filter,     &ctx->de_kidx,  &ctx->de_tiles,
                     &ctx->de_kparam,  &ctx->de_d2max};
   for (DevBuf *b : bufs)
      if (b->p) cudaFree(b->p);
   if (ctx->d_badvals) cudaFree(ctx->d_badvals);
   for (auto &p : ctx->ev_pool) {
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
   }
   for (int i = 0; i < F3B_SIDE_STREAMS; i++) {
      if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
      if (ctx->join_ev[i]) cudaEventDestroy(ctx->join_ev[i]);
   }
   if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
   if (ctx->bar_ev) cudaEventDestroy(ctx->bar_ev);
   if (ctx->inputs_ev) cudaEventDestroy(ctx->inputs_ev);
   if (ctx->timer0) cudaEventDestroy(ctx->timer0);
   if (ctx->timer1) cudaEventDestroy(ctx->timer1);
   if (ctx->stream) cudaStreamDestroy(ctx->stream);
   delete ctx;
}

const char *f3b_last_error(const f3b_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int f3b_device_info(f3b_ctx *ctx, int *info) {
   if (!ctx || !info) return 1;
   info[0] = ctx->sms;
   info[1] = ctx->iter_blocks_per_sm;
   info[2] = F3B_CHAINS_PER_GROUP;
   return 0;
}

int f3b_sync(f3b_ctx *ctx) {
   CK(cudaSetDevice(ctx->device));
   CK(cudaStreamSynchronize(ctx->stream));
   return 0;
}

/* Layout of the histogram planes. Buckets are plain row-major (F3B_HIST_PITCH=n pads the
 * rows, for experiments). What matters for speed is how evenly the chaos game's reductions
 * load the L2 slices: an attractor has a few very hot buckets (test.flam3's first flame at
 * 1080p: one bucket takes 0.27% of all hits, the top 16 take 1.9%, against 0.5% for a whole
 * average slice), so the slices that happen to hold them saturate while the rest idle
 * (ncu: lts__t_tag_requests 45% avg, 81% max), and which buckets share a slice changes with
 * every layout detail (measured 21.6 to 34.2 ms for the same frame over row pitches,
 * profiles/r01_pitch_sweep.txt). While the chaos game runs the planes are therefore
 * replicated: CTA b adds into copy b % replicas, which divides every hot bucket's load over
 * `replicas` different slices; the copies are summed before anything reads them. Replicas
 * are used as long as they all stay L2-resident (F3B_HIST_REPLICAS overrides). */
static F3HistLayout hist_layout_for(int width, int height, int count_channel, size_t *elems) {
   F3HistLayout L = {width, 1u, 0, 0u, 1u, 1u};
   const char *e = getenv("F3B_HIST_PITCH");
   if (e && atoi(e) >= width) L.pitch = atoi(e);
   const size_t buckets = (size_t)L.pitch * height;
   L.replica_stride = (buckets + 4095) & ~(size_t)4095; /* whole scramble blocks (<= 2^12 entries) */
   {
      /* scrambled staging layout: blocks of 256 entries = 4 KB (F3B_HIST_SCRAMBLE=bits, 0: plain) */
      int bits = 8;
      const char *se = getenv("F3B_HIST_SCRAMBLE");
      if (se && atoi(se) >= 0 && atoi(se) <= 12) bits = atoi(se);
      L.smask = (1u << bits) - 1u;
      L.smul = 0x9E3779B1u;
      uint32_t inv = L.smul; /* Newton: inverse of an odd number modulo 2^32 */
      for (int i = 0; i < 5; i++) inv *= 2u - L.smul * inv;
      L.sinv = inv;
   }
   const size_t plane_bytes = L.replica_stride * (sizeof(float4) + (count_channel ? sizeof(float) : 0));
   /* all copies must stay in the 126 MB L2: 3 copies at 1080p (102 MB), 4 below ~26 MB */
   size_t k = (104u << 20) / (plane_bytes ? plane_bytes : 1);
   if (k > 4) k = 4;
   if (k < 1) k = 1;
   const char *r = getenv("F3B_HIST_REPLICAS");
   if (r && atoi(r) >= 1 && atoi(r) <= 16) k = (size_t)atoi(r);
   L.replicas = (uint32_t)k;
   *elems = buckets;
   return L;
}

/* the staging layout a width x height bucket grid gets (no GPU needed; for the tests):
 * out = {entries per copy, copies, scramble mask, multiplier, inverse multiplier} */
int f3b_debug_stage_layout(int width, int height, int count_channel, uint64_t out[5]) {
   if (width <= 0 || height <= 0 || !out) return 1;
   size_t elems = 0;
   const F3HistLayout L = hist_layout_for(width, height, count_channel, &elems);
   out[0] = L.replica_stride;
   out[1] = L.replicas;
   out[2] = L.smask;
   out[3] = L.smul;
   out[4] = L.sinv;
   return 0;
}

int f3b_frame_begin(f3b_ctx *ctx, int width, int height, int count_channel, int bits) {
   if (!ctx) return 1;
   if (width <= 0 || height <= 0) return fail(ctx, "bad buffer size %dx%d", width, height);
   if (bits != 32 && bits != 33 && bits != 64) return fail(ctx, "bits must be 32, 33 or 64, not %d", bits);
   CK(cudaSetDevice(ctx->device));
   const size_t n = (size_t)width * height;
   size_t nh = 0;
   const F3HistLayout lay0 = hist_layout_for(width, height, count_channel, &nh);
   const int f64 = bits == 64;
   /* Histograms that do not fit L2 are filled through binned records (f3b_iterate.cu): bands
    * of rows whose float4 tile is F3B_BIN_REGION_MB (48) big, a record pool of F3B_BIN_MEM_MB
    * (12288). Integer buckets only (a record carries whole numbers <= 255). F3B_BINNED=0/1
    * overrides the size rule. */
   F3HistLayout lay = lay0;
   bool binned = false;
   int bin_regions = 0, bin_rows = 0;
   uint32_t bin_magic = 0;
   size_t bin_cap = 0;
   {
      const double plane_mb = (double)lay0.replica_stride * (count_channel ? 20.0 : 16.0) / 1048576.0;
      const char *be = getenv("F3B_BINNED");
      bool want = !f64 && (be ? atoi(be) != 0 : plane_mb > 110.0);
      if (want) {
         const char *re = getenv("F3B_BIN_REGION_MB");
         double region_mb = re && atof(re) > 0.0 ? atof(re) : 48.0;
         int R = (int)ceil(plane_mb / region_mb);
         if (R < 1) R = 1;
         if (R > F3B_BIN_MAX_REGIONS) R = F3B_BIN_MAX_REGIONS;
         bin_rows = (height + R - 1) / R;
         bin_regions = (height + bin_rows - 1) / bin_rows;
         bin_magic = (uint32_t)((0x100000000ull + (uint64_t)bin_rows - 1) / (uint64_t)bin_rows);
         bool ok = (size_t)bin_rows * lay0.pitch <= 0xffffffull;
         for (int y = 0; y < height && ok; y++)
            ok = (uint32_t)(((uint64_t)(uint32_t)y * bin_magic) >> 32) == (uint32_t)(y / bin_rows);
         const char *me = getenv("F3B_BIN_MEM_MB");
         const double mem_mb = me && atof(me) > 0.0 ? atof(me) : 12288.0;
         bin_cap = (size_t)(mem_mb * 1048576.0 / 256.0); /* blocks of 32 records */
         if (bin_cap > 0x7ffffff0ull) bin_cap = 0x7ffffff0ull;
         if (bin_cap < (1u << 15)) ok = false;
         binned = ok;
      }
      /* the staging planes keep their full size: f3b_iterate decides per frame which path
       * fills them (the tile of the binned path is their first bin_rows rows) */
      if (binned) {
         lay.replicas = 1;
         lay.smask = 0; /* the tile of the binned scatter is addressed by the records: plain layout */
      }
      ctx->bin_forced = binned && be != nullptr;
   }
   const size_t nalloc = lay.replica_stride * lay.replicas;
   /* staging entries are indexed in 32 bits by the walkers, the hot list and the binned records
    * (2^32 buckets would be 64 GB of staging) */
   if (nalloc >= 0xffffffffull) return fail(ctx, "bucket grid %dx%d too large", width, height);
   const void *before[2] = {ctx->hist_color.p, ctx->hist_count.p};
   if (ensure(ctx, ctx->hist_color, nalloc * sizeof(float4))) return 1;
   if (ensure(ctx, ctx->accum, n * sizeof(float4))) return 1;
   if (count_channel && ensure(ctx, ctx->hist_count, nalloc * sizeof(float))) return 1;
   if (ensure(ctx, ctx->exact_color, nh * (f64 ? sizeof(double4) : sizeof(uint4)))) return 1;
   if (count_channel && ensure(ctx, ctx->exact_count, nh * (f64 ? sizeof(double) : sizeof(uint32_t)))) return 1;
   if (ensure(ctx, ctx->sat_list, F3B_SAT_LIST_CAP * sizeof(uint32_t))) return 1;
   if (ensure(ctx, ctx->hot_list, F3B_HOT_LIST_CAP * sizeof(uint32_t))) return 1;
   if (ensure(ctx, ctx->hot_bitmap, (nalloc + 31) / 32 * sizeof(uint32_t))) return 1;
   /* the staging copies are zero between a merge and the next launch: clear them only when
    * they are new, were left behind by a failed frame, or the layout changed under them */
   if (before[0] != ctx->hist_color.p || (count_channel && before[1] != ctx->hist_count.p) ||
       ctx->lay.replica_stride != lay.replica_stride || ctx->lay.replicas != lay.replicas ||
       ctx->count_channel != (count_channel ? 1 : 0))
      ctx->staging_dirty = true;
   if (binned) {
      if (ensure(ctx, ctx->bin_records, bin_cap * 32 * sizeof(uint2))) return 1;
      if (ensure(ctx, ctx->bin_table, bin_cap)) return 1;
      if (ensure(ctx, ctx->bin_pos, 4 * sizeof(unsigned int))) return 1;
   }
   ctx->binned = binned;
   ctx->bin_regions = bin_regions;
   ctx->bin_region_rows = bin_rows;
   ctx->bin_magic = bin_magic;
   ctx->bin_cap = bin_cap;
   ctx->bin_tile_elems = binned ? (size_t)bin_rows * lay.pitch : 0;
   ctx->lay = lay;
   ctx->hist_elems = nh;
   ctx->width = width;
   ctx->height = height;
   ctx->count_channel = count_channel ? 1 : 0;
   ctx->f64 = f64;
   ctx->hot_total = 0.0;
   ctx->hot_measured = false;
   ctx->hot_from_merge = false;
   CK(cudaMemsetAsync(ctx->d_badvals, 0, F3B_NCOUNTERS * sizeof(unsigned long long), ctx->stream));
   if (ctx->staging_dirty) {
      CK(cudaMemsetAsync(ctx->hist_color.p, 0, nalloc * sizeof(float4), ctx->stream));
      if (count_channel) CK(cudaMemsetAsync(ctx->hist_count.p, 0, nalloc * sizeof(float), ctx->stream));
      ctx->staging_dirty = false;
   }
   if (f3b_hist_clear(ctx)) return 1;
   return f3b_accum_clear(ctx);
}

int f3b_hist_clear(f3b_ctx *ctx) {
   const size_t n = ctx->hist_elems;
   CK(cudaSetDevice(ctx->device));
   CK(cudaMemsetAsync(ctx->exact_color.p, 0, n * (ctx->f64 ? sizeof(double4) : sizeof(uint4)), ctx->stream));
   if (ctx->count_channel)
      CK(cudaMemsetAsync(ctx->exact_count.p, 0, n * (ctx->f64 ? sizeof(double) : sizeof(uint32_t)), ctx->stream));
   if (ctx->swept) /* the saturation list belongs to the histogram being discarded */
      CK(cudaMemsetAsync(ctx->d_badvals + F3B_CTR_LOST + 1, 0, sizeof(unsigned long long), ctx->stream));
   ctx->swept = false;
   ctx->view_reduced = false; /* the post-pass reads this context's own buckets again */
   return 0;
}

/* The clear is deferred: the first density pass of a frame writes every accumulator, so the
 * memset (and that pass's read of the zeros) is only issued when something else looks first. */
int f3b_accum_clear(f3b_ctx *ctx) {
   ctx->accum_zero_pending = true;
   return 0;
}

static int accum_materialize(f3b_ctx *ctx) {
   if (!ctx->accum_zero_pending) return 0;
   const size_t n = (size_t)ctx->width * ctx->height;
   CK(cudaSetDevice(ctx->device));
   CK(cudaMemsetAsync(ctx->accum.p, 0, n * sizeof(float4), ctx->stream));
   ctx->accum_zero_pending = false;
   return 0;
}

/* Chain planning. The reference gives each of its `nthreads` workers batch_size/nthreads
 * iterations and cuts them into sub-batches of at most sub_batch_size (rect.c:840, 279-284);
 * a sub-batch is a chain: random start, `fuse` discarded iterations, then stored ones. So
 * the reference runs  nthreads * ceil((n / nthreads) / sub_batch_size)  chains, and that is
 * the MINIMUM used here, with `workers` = the caller's nthreads.
 *
 * Chain length is an observable: only `fuse` (15) iterations are discarded after a restart,
 * so for slowly converging systems the image depends on how often the orbit restarts
 * (measured: the reference's own output moves by 10+ dB of PSNR between sub_batch_size 10000
 * and 2048 on such genomes, tests/test_gpu_variations.py). Beyond the minimum the count is
 * therefore only rounded up, never by more than 2.5x (chains stay >= 0.4 sub_batch_size
 * unless the caller's own nthreads says otherwise):
 *   - to whole rounds of a node's worth of resident threads (F3B_PLAN_GPUS GPUs), when the
 *     minimum is within 2.5x of that: the plan - and with it the image for a given seed -
 *     then does not depend on how many GPUs of the node share the frame, and each of them
 *     gets whole rounds of its resident threads (at 8 GPUs a 1080p q2000 frame would
 *     otherwise leave every GPU half empty and latency bound);
 *   - else to whole rounds of one GPU's resident threads when there are more chains than
 *     that, so no tail wave is left. */
#define F3B_PLAN_GPUS 8
static uint32_t plan_chains(f3b_ctx *ctx, int blocks_per_sm, uint64_t n_iters, uint32_t max_chain_len,
                            uint32_t n_jobs, uint32_t workers);
uint32_t f3b_plan_chains(f3b_ctx *ctx, uint64_t n_iters, uint32_t max_chain_len, uint32_t n_jobs,
                         uint32_t workers) {
   if (!ctx) return 0;
   return plan_chains(ctx, ctx->iter_blocks_per_sm, n_iters, max_chain_len, n_jobs, workers);
}
/* blocks_per_sm: of the kernel class that will run the chains (the SIMPLE kernels keep 3 CTAs
 * per SM resident, the generic ones 2) */
static uint32_t plan_chains(f3b_ctx *ctx, int blocks_per_sm, uint64_t n_iters, uint32_t max_chain_len,
                            uint32_t n_jobs, uint32_t workers) {
   if (!ctx || n_iters == 0) return 0;
   if (max_chain_len == 0) max_chain_len = 10000;
   if (n_jobs == 0) n_jobs = 1;
   if (workers == 0) workers = 1;
   const uint64_t T = (uint64_t)ctx->sms * blocks_per_sm * F3B_CHAINS_PER_GROUP;
   uint64_t Ts = T / n_jobs;
   if (Ts < 32) Ts = 32;
   if (Ts >= F3B_CHAINS_PER_GROUP) Ts -= Ts % F3B_CHAINS_PER_GROUP;
   const uint64_t Tv = Ts * F3B_PLAN_GPUS;

   const uint64_t per_worker = (n_iters + workers - 1) / workers;
   uint64_t chains = (uint64_t)workers * ((per_worker + max_chain_len - 1) / max_chain_len);
   if (chains > n_iters) chains = n_iters;
   if (chains == 0) chains = 1;
   if (2 * Tv <= 5 * chains)
      chains = ((chains + Tv - 1) / Tv) * Tv;
   else if (chains > Ts)
      chains = ((chains + Ts - 1) / Ts) * Ts;
   if (chains > 0xfffffff0ull) chains = 0xfffffff0ull;
   return (uint32_t)chains;
}

/* What stays the same between the frames of an animation or between repeated renders of a
 * genome: its structure (xform count, xaos, final xform, each xform's flags and variation
 * ids - not the coefficients) and the frame size. Keys the hot-bucket feedback below. */
static uint64_t genome_signature(const uint32_t *w0, int width, int height) {
   uint64_t sig = 0xcbf29ce484222325ull ^ ((uint64_t)width << 32) ^ (uint64_t)height;
   const uint32_t nx = w0[F3B_H_NXF];
   sig = (sig ^ nx ^ ((uint64_t)w0[F3B_H_XAOS] << 8) ^ ((uint64_t)w0[F3B_H_FINAL] << 16)) * 0x100000001b3ull;
   for (uint32_t q = 0; q < nx; q++) {
      const uint32_t rec = w0[F3B_H_XF_BASE] + q * w0[F3B_H_XF_STRIDE];
      sig = (sig ^ w0[rec + F3B_X_FLAGS]) * 0x100000001b3ull;
      for (uint32_t pv = rec + F3B_X_VARS; pv < rec + F3B_X_VARS + (w0[rec + F3B_X_FLAGS] >> 8);) {
         sig = (sig ^ (w0[pv] & 0xffu)) * 0x100000001b3ull;
         pv += 2 + (w0[pv] >> 8);
      }
   }
   return sig;
}

void *f3b_pinned_alloc(size_t bytes) {
   void *p = nullptr;
   /* portable: a frame spread over several GPUs uploads the same packed samples to each */
   if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
   }
   return p;
}
void f3b_pinned_free(void *p) {
   if (p) cudaFreeHost(p);
}
int f3b_inputs_consumed(f3b_ctx *ctx) {
   if (!ctx || !ctx->inputs_pending) return 0;
   CK(cudaSetDevice(ctx->device));
   CK(cudaEventSynchronize(ctx->inputs_ev));
   ctx->inputs_pending = false;
   return 0;
}

/* every sample qualifies for the SIMPLE kernels (<= 4 xforms, no xaos, no final xform, step
 * palette, all xforms opaque)? */
static bool blob_is_simple(const uint32_t *w) {
   return !w[F3B_H_XAOS] && w[F3B_H_CHOOSER_STRIDE] == 4 && (w[F3B_H_PLOT] & F3B_PLOT_READY) &&
          !w[F3B_H_PALMODE] && (int)w[F3B_H_FINAL] < 0;
}

int f3b_iterate(f3b_ctx *ctx, const void *blobs, size_t blob_bytes, f3b_job *jobs, int n_jobs,
                uint64_t seed, int floor_adds, uint32_t max_chain_len, uint32_t workers, int rank, int nranks) {
   if (!ctx || !blobs || !jobs || n_jobs <= 0) return fail(ctx, "f3b_iterate: bad arguments");
   if (nranks <= 0 || rank < 0 || rank >= nranks) return fail(ctx, "f3b_iterate: bad rank");
   if (!ctx->hist_color.p || !ctx->exact_color.p) return fail(ctx, "f3b_iterate: f3b_frame_begin not called");
   CK(cudaSetDevice(ctx->device));

   /* plan chains and groups */
   std::vector<uint32_t> group_job;
   size_t max_blob = 0;
   uint64_t total_chains = 0;
   /* which kernel class runs this batch decides how many threads are resident, which the
    * chain plan rounds to */
   bool all_simple = getenv("F3B_NO_SIMPLE") == nullptr;
   for (int j = 0; j < n_jobs && all_simple; j++) {
      if (jobs[j].blob_offset % 16 || jobs[j].blob_offset + F3B_H_SIZE * 4 > blob_bytes)
         return fail(ctx, "f3b_iterate: job %d blob offset invalid", j);
      all_simple = blob_is_simple((const uint32_t *)((const char *)blobs + jobs[j].blob_offset));
   }
   const int plan_bps = all_simple ? ctx->iter_blocks_per_sm : ctx->iter_blocks_per_sm_generic;
   for (int j = 0; j < n_jobs; j++) {
      f3b_job &jb = jobs[j];
      if (jb.n_chains == 0) jb.n_chains = plan_chains(ctx, plan_bps, jb.n_iters, max_chain_len, n_jobs, workers);
      total_chains += jb.n_chains;
   }
   /* Chains per group (= threads per CTA): 256 when there is plenty of work; small batches
    * (draft quality, or the 1000 short temporal samples of an animation frame, where a
    * sample may have a single chain) use narrower groups so that the groups of all samples
    * are resident together and the serial walks overlap instead of queueing. */
   int group_size = F3B_CHAINS_PER_GROUP;
   for (; group_size > 32; group_size /= 2) {
      uint64_t groups = 0;
      for (int j = 0; j < n_jobs; j++) groups += (jobs[j].n_chains + group_size - 1) / group_size;
      const bool fills_gpu = groups / (uint64_t)nranks >= 3ull * (uint64_t)ctx->sms; /* this rank's share */
      const bool lanes_used = 2 * total_chains >= groups * (uint64_t)group_size;
      if (fills_gpu && lanes_used) break;
   }
   for (int j = 0; j < n_jobs; j++) {
      f3b_job &jb = jobs[j];
      if (jb.blob_offset % 16 || jb.blob_offset + F3B_H_SIZE * 4 > blob_bytes)
         return fail(ctx, "f3b_iterate: job %d blob offset invalid", j);
      const uint32_t words = ((const uint32_t *)((const char *)blobs + jb.blob_offset))[F3B_H_WORDS];
      if (words < F3B_H_SIZE || words > F3B_MAX_BLOB_WORDS ||
          jb.blob_offset + (size_t)words * 4 > blob_bytes)
         return fail(ctx, "f3b_iterate: job %d blob of %u words does not fit (max %u)", j, words,
                     F3B_MAX_BLOB_WORDS);
      if ((size_t)words * 4 > max_blob) max_blob = (size_t)words * 4;
      jb.first_group = (uint32_t)group_job.size();
      const uint32_t ng = (jb.n_chains + (uint32_t)group_size - 1) / (uint32_t)group_size;
      for (uint32_t g = 0; g < ng; g++) group_job.push_back((uint32_t)j);
   }
   const uint32_t n_groups = (uint32_t)group_job.size();
   if (n_groups == 0) return 0;

   if (ensure(ctx, ctx->blobs, blob_bytes)) return 1;
   if (ensure(ctx, ctx->jobs, sizeof(f3b_job) * (size_t)n_jobs)) return 1;
   if (ensure(ctx, ctx->groups, sizeof(uint32_t) * (size_t)n_groups)) return 1;
   CK(cudaMemcpyAsync(ctx->blobs.p, blobs, blob_bytes, cudaMemcpyHostToDevice, ctx->stream));
   CK(cudaMemcpyAsync(ctx->jobs.p, jobs, sizeof(f3b_job) * (size_t)n_jobs, cudaMemcpyHostToDevice,
                      ctx->stream));
   CK(cudaMemcpyAsync(ctx->groups.p, group_job.data(), sizeof(uint32_t) * (size_t)n_groups,
                      cudaMemcpyHostToDevice, ctx->stream));
   ctx->h2d_bytes += blob_bytes + sizeof(f3b_job) * (size_t)n_jobs + sizeof(uint32_t) * (size_t)n_groups;
   {
      /* jobs and groups are small and pageable: their copies are staged before the call
       * returns. A page-locked blob buffer is read by DMA later: mark when it is free again;
       * a pageable one is waited for here. */
      cudaPointerAttributes pa;
      const bool pinned = cudaPointerGetAttributes(&pa, blobs) == cudaSuccess && pa.type == cudaMemoryTypeHost;
      cudaGetLastError(); /* unregistered pageable memory may report an error on old drivers */
      if (!ctx->inputs_ev) CK(cudaEventCreateWithFlags(&ctx->inputs_ev, cudaEventDisableTiming));
      CK(cudaEventRecord(ctx->inputs_ev, ctx->stream));
      ctx->inputs_pending = true;
      if (!pinned && f3b_inputs_consumed(ctx)) return 1;
   }

   F3IterateParams P;
   P.blobs = (const uint8_t *)ctx->blobs.p;
   P.jobs = (const f3b_sample_job *)ctx->jobs.p;
   P.group_job = (const uint32_t *)ctx->groups.p;
   P.n_groups = n_groups;
   P.rank = (uint32_t)rank;
   P.nranks = (uint32_t)nranks;
   P.hist_color = (float4 *)ctx->hist_color.p;
   P.hist_count = ctx->count_channel ? (float *)ctx->hist_count.p : nullptr;
   P.width = ctx->width;
   P.height = ctx->height;
   P.lay = ctx->lay;
   P.floor_adds = floor_adds;
   P.seed = seed;
   P.badvals = ctx->d_badvals;
   /* Binned scatter or direct reductions? (only a question when the planes exceed L2.) The
    * tile of the binned path collects a whole batch - up to the record pool's worth of
    * iterations - before it is merged into the integer buckets, so its fullest bucket must
    * stay below 2^24 over a batch: known only from the previous frame of the same genome
    * structure and size (the merge records the fullest bucket's share). Spread-out histograms
    * (motion blur over many temporal samples) qualify and gain ~20 %; a still with hot buckets
    * is both exact and faster through the direct path with its sweepers. */
   bool binned = false;
   double bin_safe_iters = 0.0;
   if (ctx->binned && floor_adds) {
      const uint64_t sig = genome_signature((const uint32_t *)((const char *)blobs + jobs[0].blob_offset),
                                            ctx->width, ctx->height);
      const bool known = ctx->hot_share >= 0.0 && ctx->hot_share_signature == sig;
      /* iterations a batch may hold so that the fullest bucket gains < 2^23 in it */
      const double safe_iters = known ? 8388608.0 / (255.0 * (ctx->hot_share > 1e-12 ? ctx->hot_share : 1e-12)) : 0.0;
      binned = ctx->bin_forced || safe_iters >= 2e8; /* smaller batches are latency bound */
      bin_safe_iters = ctx->bin_forced ? 1e18 : safe_iters;
   }
   P.exact_color = ctx->exact_color.p;
   P.exact_count = ctx->count_channel ? ctx->exact_count.p : nullptr;
   P.exact_f64 = (uint32_t)ctx->f64;
   P.sweep_state = (unsigned int *)(ctx->d_badvals + F3B_CTR_TICKET);
   P.sweep_hot_list = (uint32_t *)ctx->hot_list.p;
   P.sweep_hot_bitmap = (unsigned int *)ctx->hot_bitmap.p;
   P.exact_flags = ctx->d_badvals + F3B_CTR_LOST;
   P.sat_list = (uint32_t *)ctx->sat_list.p;
   P.chain_threads = (uint32_t)group_size;
   {
      /* Sweepers (f3b_iterate.cu) run unless no staged sum can reach the carry threshold at
       * all. After the first passes a cold pass over all staging entries is due once per
       * sweep_gap: 2 ms, or longer for big histograms so that the passes stay below
       * ~100 GB/s of reads. F3B_SWEEP=0/1, F3B_SWEEP_GAP_US override. */
      uint64_t it = 0;
      for (int j = 0; j < n_jobs; j++) it += jobs[j].n_iters;
      double carry = ctx->f64 ? (double)F3B_CARRY_F64 : (double)F3B_CARRY_U32;
      const char *ce = getenv("F3B_CARRY"); /* test hook: a lower threshold exercises the carries */
      if (ce && atof(ce) >= 1.0 && atof(ce) < carry) carry = atof(ce);
      P.carry_thresh = (float)carry;
      P.sweep_on = ((double)it / nranks) * 255.0 >= carry ? 1u : 0u;
      const char *e = getenv("F3B_SWEEP");
      if (e) P.sweep_on = atoi(e) ? 1u : 0u;
      if (binned) P.sweep_on = 0; /* the tile is merged into the exact buckets after every batch */
      const double plane_bytes = (double)(ctx->lay.replica_stride * ctx->lay.replicas) * (ctx->count_channel ? 20.0 : 16.0);
      double gap_us = plane_bytes / 100e9 * 1e6;
      if (gap_us < 2000.0) gap_us = 2000.0;
      const char *ge = getenv("F3B_SWEEP_GAP_US");
      if (ge && atof(ge) > 0.0) gap_us = atof(ge);
      P.sweep_gap_ns = gap_us * 1e3 > 4e9 ? 4000000000u : (uint32_t)(gap_us * 1e3);
      const char *me = getenv("F3B_SWEEP_MODE");
      /* histograms beyond L2: the cold passes are slow beside the walkers' DRAM scatter, the
       * walkers report a sample of the entries they add to (kernel MODE 2) */
      P.sweep_mode = me ? (uint32_t)atoi(me) : (plane_bytes > 112e6 ? 11u : 3u);
      if (P.sweep_on) {
         CK(cudaMemsetAsync(P.sweep_state, 0, 4 * sizeof(unsigned int), ctx->stream));
         CK(cudaMemsetAsync(ctx->hot_bitmap.p, 0,
                            (ctx->lay.replica_stride * ctx->lay.replicas + 31) / 32 * sizeof(uint32_t), ctx->stream));
      }
   }
   /* binned scatter: per-warp region queues behind the blob in shared memory */
   P.bin_regions = 0;
   size_t smem_bytes = max_blob;
   if (binned) {
      P.bin_regions = (uint32_t)ctx->bin_regions;
      P.bin_region_rows = (uint32_t)ctx->bin_region_rows;
      P.bin_row_magic = ctx->bin_magic;
      P.bin_cap = (uint32_t)ctx->bin_cap;
      P.bin_smem_off = (uint32_t)((max_blob + 15) / 16 * 4);
      P.bin_warp_words = (uint32_t)(32 + ctx->bin_regions * 130); /* 65 records per region: queues start in different banks */
      P.bin_records = (uint2 *)ctx->bin_records.p;
      P.bin_block_region = (uint8_t *)ctx->bin_table.p;
      P.bin_pos = (unsigned int *)ctx->bin_pos.p;
      smem_bytes = (size_t)P.bin_smem_off * 4 + (size_t)(group_size / 32) * P.bin_warp_words * 4;
   }
   /* warp-uniform xform step (f3b_iterate.cu): CTAs of >= 4 warps, not with the binned
    * scatter; per sample it is off for xaos genomes. It pays when the xforms run DIFFERENT
    * variation code, i.e. when lanes that picked different xforms would diverge (random
    * genomes with several variations per xform: 1.6-5.4x, profiles/r02_warp_uniform_ab.txt); a
    * classic flame whose xforms share their code path loses 12 % to the CTA-wide exchange of
    * the walkers. So: on when the standard xforms of the first sample differ in their
    * variation lists or flags. F3B_WU=0/1 overrides. */
   P.warp_uniform = 0;
   P.wu_smem_off = 0;
   {
      const uint32_t *w0 = (const uint32_t *)((const char *)blobs + jobs[0].blob_offset);
      bool heterogeneous = false;
      uint64_t first_sig = 0;
      for (uint32_t q = 0; q < w0[F3B_H_NSTD]; q++) {
         const uint32_t rec = w0[F3B_H_XF_BASE] + q * w0[F3B_H_XF_STRIDE];
         uint64_t sig = 0xcbf29ce484222325ull ^ (w0[rec + F3B_X_FLAGS] & 0xffu);
         for (uint32_t pv = rec + F3B_X_VARS; pv < rec + F3B_X_VARS + (w0[rec + F3B_X_FLAGS] >> 8);) {
            sig = (sig ^ (w0[pv] & 0xffu)) * 0x100000001b3ull;
            pv += 2 + (w0[pv] >> 8);
         }
         if (q == 0) first_sig = sig;
         else if (sig != first_sig) heterogeneous = true;
      }
      const char *we = getenv("F3B_WU");
      /* (xforms that differ only in their linear / spherical weights - the fixed-function slots
       * every xform runs - have no code to diverge in: the classic flames only lose, -12 %) */
      const bool want = we ? atoi(we) != 0 : heterogeneous;
      if (want && !binned && group_size >= 128) {
         P.warp_uniform = 1;
         P.wu_smem_off = (uint32_t)((smem_bytes + 15) / 16 * 4);
         smem_bytes = (size_t)P.wu_smem_off * 4 + (size_t)group_size * 3 * sizeof(float);
      }
   }
   const int bps_simple = f3b_iterate_blocks_per_sm(group_size, smem_bytes, true);
   const int bps_generic = f3b_iterate_blocks_per_sm(group_size, smem_bytes, false);
   if (bps_simple <= 0 || bps_generic <= 0)
      return fail(ctx, "f3b_iterate: kernel does not fit with %zu bytes of shared memory", smem_bytes);
   ctx->staging_dirty = true;

   /* One launch per `cap` temporal samples (their warp-uniform camera and chooser constants
    * travel through the constant cache); a still frame is one launch. */
   const int cap = f3b_iterate_const_capacity();
   std::vector<F3ConstSample> consts((size_t)(n_jobs < cap ? n_jobs : cap));
   int launches = 0;
   if (stage_begin(ctx, F3B_STAGE_ITERATE)) return 1;
   P.probe_iters = 0;
   /* device counters: [0] bad values, [1] fullest bucket of the frame (float bits, written by
    * the density pass), [2] lanes adding / [3] lanes sharing a bucket in the probe, [4] the
    * probe's own bad values (discarded) */
   P.probe_counts = ctx->d_badvals + 2;
   bool aggregate = false;
   /* several launches (more temporal samples than fit the parameter space) are independent:
    * they go to side streams so that their CTAs can be resident together */
   const bool fork = n_jobs > cap;
   if (fork) {
      for (int i = 0; i < F3B_SIDE_STREAMS; i++)
         if (!ctx->side[i]) CK(cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking));
      if (!ctx->fork_ev) CK(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
      for (int i = 0; i < F3B_SIDE_STREAMS; i++)
         if (!ctx->join_ev[i]) CK(cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming));
      CK(cudaEventRecord(ctx->fork_ev, ctx->stream));
      for (int i = 0; i < F3B_SIDE_STREAMS; i++) CK(cudaStreamWaitEvent(ctx->side[i], ctx->fork_ev, 0));
   }
   int chunk = 0;
   for (int j0 = 0; j0 < n_jobs; j0 += cap, chunk++) {
      const int j1 = (j0 + cap < n_jobs) ? j0 + cap : n_jobs;
      /* every sample of the launch qualifies for the SIMPLE kernel (F3B_NO_SIMPLE=1 forces
       * the generic kernel, for A/B timing) */
      bool simple = getenv("F3B_NO_SIMPLE") == nullptr;
      for (int j = j0; j < j1; j++) {
         const uint32_t *w = (const uint32_t *)((const char *)blobs + jobs[j].blob_offset);
         const float *f = (const float *)w;
         F3ConstSample &c = consts[(size_t)(j - j0)];
         c.bounds = make_float4(f[F3B_H_B0], f[F3B_H_B1], f[F3B_H_B2], f[F3B_H_B3]);
         c.cs = make_float4(f[F3B_H_WS0], f[F3B_H_HS1], f[F3B_H_RCX], f[F3B_H_RCY]);
         c.rot = make_float4(f[F3B_H_ROT00], f[F3B_H_ROT01], f[F3B_H_ROT10], f[F3B_H_ROT11]);
         const uint32_t *cum = w + w[F3B_H_OFF_CHOOSER];
         c.cum = make_uint4(cum[0], cum[1], cum[2], cum[3]);
         bool any_post = false;
         for (uint32_t q = 0; q < w[F3B_H_NXF]; q++)
            if (w[w[F3B_H_XF_BASE] + q * w[F3B_H_XF_STRIDE] + F3B_X_FLAGS] & F3B_XF_HAS_POST) any_post = true;
         const uint32_t hf =
             (any_post ? F3B_HF_ANY_POST : 0u) | (w[F3B_H_XAOS] ? F3B_HF_XAOS : 0u) | (w[F3B_H_ROTATE] ? F3B_HF_ROTATE : 0u) |
             ((w[F3B_H_PLOT] & F3B_PLOT_READY) ? F3B_HF_PLOT_READY : 0u) |
             ((w[F3B_H_PLOT] & F3B_PLOT_PACKED) ? F3B_HF_PAL_PACKED : 0u) |
             (w[F3B_H_PALMODE] ? F3B_HF_PAL_LINEAR : 0u) | ((int)w[F3B_H_FINAL] >= 0 ? F3B_HF_FINAL : 0u) |
             ((!w[F3B_H_XAOS] && w[F3B_H_CHOOSER_STRIDE] == 4) ? F3B_HF_CHOOSER_CONST : 0u) |
             (w[F3B_H_CHOOSER_STRIDE] << 8) | (w[F3B_H_CMAPSIZE] << 20);
         const uint32_t final_off =
             (int)w[F3B_H_FINAL] >= 0 ? w[F3B_H_XF_BASE] + w[F3B_H_FINAL] * w[F3B_H_XF_STRIDE] : 0u;
         c.lay = make_uint4(hf, w[F3B_H_XF_BASE] | (w[F3B_H_XF_STRIDE] << 16),
                            w[F3B_H_OFF_PALETTE] | (w[F3B_H_OFF_CHOOSER] << 16), final_off);
         /* packed pick table of the SIMPLE kernel: the walk takes the first entry whose
          * (count << 18 | offset) exceeds (k << 18 | 0x3ffff), else the last */
         uint32_t pk[4];
         const uint32_t nstd = w[F3B_H_NSTD];
         for (uint32_t q = 0; q < 4; q++) {
            const uint32_t xi = (nstd == 0) ? 0 : (q < nstd ? q : nstd - 1);
            const uint32_t off = w[F3B_H_XF_BASE] + xi * w[F3B_H_XF_STRIDE];
            const uint32_t cnt = cum[q < nstd ? q : (nstd ? nstd - 1 : 0)];
            pk[q] = (cnt >= 16384u ? 0xfffc0000u : (cnt << 18)) | off;
         }
         /* k = 16383 compares below no entry (a count of 16384 saturates to the same 14 bits)
          * and falls through to the last slot: that slot must be the FIRST xform whose
          * cumulative count reaches 16384, which is what the reference's table holds at
          * [16383] when trailing xforms have zero weight (flam3.c:196-216). Every smaller k
          * already stops at or before that xform's own slot. */
         for (uint32_t q = 0; q < 4 && q < nstd; q++)
            if (cum[q] >= 16384u) {
               pk[3] = 0xfffc0000u | (w[F3B_H_XF_BASE] + q * w[F3B_H_XF_STRIDE]);
               break;
            }
         c.pick = make_uint4(pk[0], pk[1], pk[2], pk[3]);
         if ((hf & (F3B_HF_CHOOSER_CONST | F3B_HF_PLOT_READY | F3B_HF_FINAL | F3B_HF_XAOS |
                    F3B_HF_PAL_LINEAR)) != (F3B_HF_CHOOSER_CONST | F3B_HF_PLOT_READY))
            simple = false;
      }
      const int bps = simple ? bps_simple : bps_generic;
      P.job_base = (uint32_t)j0;
      P.group_begin = jobs[j0].first_group;
      P.n_groups = (j1 < n_jobs) ? jobs[j1].first_group : n_groups;
      const uint32_t span = P.n_groups - P.group_begin;
      const uint32_t local_groups = (span + (uint32_t)nranks - 1) / (uint32_t)nranks;
      if (span == 0) continue;
      int grid = ctx->sms * bps;
      if ((uint32_t)grid > local_groups) grid = (int)local_groups;
      cudaStream_t st = fork ? ctx->side[chunk % F3B_SIDE_STREAMS] : ctx->stream;
      if (binned) {
         /* A BATCH is as many of this rank's chain groups as the record pool holds for certain
          * (one record per iteration in the worst case). Its launches - one per group of
          * temporal samples that share a parameter block - run side by side on the side
          * streams; then the pool is applied region by region to the L2-resident tile and the
          * tile merged into the region's exact buckets; repeat. */
         if (chunk == 0) {
            uint64_t total_iters = 0;
            for (int j = 0; j < n_jobs; j++) total_iters += jobs[j].n_iters;
            ctx->hot_signature = genome_signature(
                (const uint32_t *)((const char *)blobs + jobs[0].blob_offset), ctx->width, ctx->height);
            ctx->hot_iters = total_iters;
         }
         /* pool blocks that may stay unfilled: the last hand-out of 32 and one padded block per
          * region for every warp of every launch that can be in flight */
         const uint64_t slack = (uint64_t)(F3B_SIDE_STREAMS + 1) * ctx->sms * bps * (group_size / 32) *
                                (uint64_t)(ctx->bin_regions + 33);
         if (ctx->bin_cap <= slack + 1024) return fail(ctx, "f3b_iterate: record pool too small; raise F3B_BIN_MEM_MB");
         uint64_t room = (ctx->bin_cap - slack) * 32;
         if ((double)room > bin_safe_iters) room = (uint64_t)bin_safe_iters;
         auto batch_begin = [&]() -> int {
            CK(cudaMemsetAsync(ctx->bin_pos.p, 0, 4 * sizeof(unsigned int), ctx->stream));
            CK(cudaMemsetAsync(ctx->bin_table.p, 0xff, ctx->bin_cap, ctx->stream));
            if (fork) {
               CK(cudaEventRecord(ctx->fork_ev, ctx->stream));
               for (int i = 0; i < F3B_SIDE_STREAMS; i++) CK(cudaStreamWaitEvent(ctx->side[i], ctx->fork_ev, 0));
            }
            ctx->bin_batch_open = true;
            ctx->bin_batch_iters = 0;
            return 0;
         };
         auto batch_end = [&]() -> int {
            if (fork)
               for (int i = 0; i < F3B_SIDE_STREAMS; i++) {
                  CK(cudaEventRecord(ctx->join_ev[i], ctx->side[i]));
                  CK(cudaStreamWaitEvent(ctx->stream, ctx->join_ev[i], 0));
               }
            for (int r = 0; r < ctx->bin_regions; r++) {
               if (!f3b_launch_bin_apply((const uint2 *)ctx->bin_records.p, (const uint8_t *)ctx->bin_table.p,
                                         (const unsigned int *)ctx->bin_pos.p, (uint32_t)ctx->bin_cap, (uint32_t)r,
                                         (float4 *)ctx->hist_color.p,
                                         ctx->count_channel ? (float *)ctx->hist_count.p : nullptr, ctx->stream))
                  return fail(ctx, "record apply launch failed: %s", cudaGetErrorString(cudaGetLastError()));
               const size_t off = (size_t)r * ctx->bin_tile_elems;
               F3MergeParams M;
               M.color = (float4 *)ctx->hist_color.p;
               M.count = ctx->count_channel ? (float *)ctx->hist_count.p : nullptr;
               M.exact_color = (char *)ctx->exact_color.p + off * sizeof(uint4);
               M.exact_count = ctx->count_channel ? (char *)ctx->exact_count.p + off * sizeof(uint32_t) : nullptr;
               M.f64 = 0;
               M.elems = ctx->hist_elems - off < ctx->bin_tile_elems ? ctx->hist_elems - off : ctx->bin_tile_elems;
               M.lay = ctx->lay;
               M.hot_max = (unsigned int *)(ctx->d_badvals + 1);
               M.exact_flags = ctx->d_badvals + F3B_CTR_LOST;
               M.sat_list = (const uint32_t *)ctx->sat_list.p;
               M.sat_check = 0;
               if (f3b_launch_hist_merge(M, ctx->stream) < 0)
                  return fail(ctx, "tile merge launch failed: %s", cudaGetErrorString(cudaGetLastError()));
               launches += 2;
            }
            ctx->bin_batch_open = false;
            return 0;
         };
         auto launch_range = [&](uint32_t ga, uint32_t gb, uint32_t mine) -> int {
            F3IterateParams Q = P;
            Q.group_begin = ga;
            Q.n_groups = gb;
            int qgrid = ctx->sms * bps;
            if ((uint32_t)qgrid > mine) qgrid = (int)mine;
            cudaStream_t qs = fork ? ctx->side[(ctx->bin_launch_seq++) % F3B_SIDE_STREAMS] : ctx->stream;
            if (!f3b_launch_iterate(Q, consts.data(), j1 - j0, ctx->count_channel != 0, simple, true, qgrid,
                                    group_size, smem_bytes, qs))
               return fail(ctx, "chaos-game launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            launches++;
            return 0;
         };
         if (chunk == 0) ctx->bin_batch_open = false;
         const uint32_t g_first = P.group_begin + ((uint32_t)rank + (uint32_t)nranks - P.group_begin % (uint32_t)nranks) % (uint32_t)nranks;
         const uint32_t g_end = P.n_groups;
         uint32_t ga = P.group_begin, mine = 0;
         for (uint32_t g = g_first; g < g_end; g += (uint32_t)nranks) {
            const f3b_job &jb = jobs[group_job[g]];
            const uint64_t first_chain = (uint64_t)(g - jb.first_group) * group_size;
            uint64_t nch = jb.n_chains > first_chain ? jb.n_chains - first_chain : 0;
            if (nch > (uint64_t)group_size) nch = group_size;
            const uint64_t gi = nch * (jb.n_iters / jb.n_chains + 1);
            if (gi > room)
               return fail(ctx, "f3b_iterate: one chain group (%llu iterations) exceeds the record pool (%llu records); raise F3B_BIN_MEM_MB",
                           (unsigned long long)gi, (unsigned long long)room);
            if (ctx->bin_batch_open && ctx->bin_batch_iters + gi > room) { /* the pool is full: close the batch */
               if (mine && launch_range(ga, g, mine)) return 1;
               if (batch_end()) return 1;
               ga = g;
               mine = 0;
            }
            if (!ctx->bin_batch_open && batch_begin()) return 1;
            ctx->bin_batch_iters += gi;
            mine++;
         }
         if (mine && launch_range(ga, g_end, mine)) return 1;
         if (j1 >= n_jobs && ctx->bin_batch_open && batch_end()) return 1; /* last parameter block */
         continue;
      }
      if (chunk == 0) {
         /* Does the attractor collapse onto a few buckets? Then the adds must be aggregated
          * across the warp, or they serialise on a handful of L2 addresses. A spread-out flame
          * pays ~5% for the aggregation check, so it is decided by measurement: the first
          * 64 iterations of one group per SM run once without adding anything and count how
          * many lanes share their bucket with another lane (~0.1 ms; frames too small to
          * care are not probed). When the previous frame had the same genome structure and
          * size, the merge of its histogram copies (or its whole-frame density pass) has
          * measured the share of the fullest bucket, and a clear-cut share decides without a
          * probe. F3B_AGG=0/1 overrides. */
         const char *e = getenv("F3B_AGG");
         uint64_t total_iters = 0;
         for (int j = 0; j < n_jobs; j++) total_iters += jobs[j].n_iters;
         const uint64_t sig = genome_signature(
             (const uint32_t *)((const char *)blobs + jobs[0].blob_offset), ctx->width, ctx->height);
         ctx->hot_signature = sig;
         ctx->hot_iters = total_iters;
         ctx->hot_total = 255.0 * (double)total_iters; /* of one batch, like the bucket maximum */
         const bool known = ctx->hot_share >= 0.0 && ctx->hot_share_signature == sig;
         if (e) aggregate = atoi(e) != 0;
         /* one bucket with share s of the hits gives a lane a duplicate among its 31 peers with
          * probability ~31 s: the probe's 1/8 criterion corresponds to s ~ 0.4% */
         else if (known && ctx->hot_share < 0.0035) aggregate = false; /* spread out last time */
         else if (known && ctx->hot_share > 0.02) aggregate = true;   /* one bucket took > 2% */
         else if (total_iters / (uint64_t)nranks >= (P.sweep_on ? 1000000ull : 50000000ull)) {
            /* (with integer buckets the answer also decides whether hot buckets get their
             * aggregated sums added exactly, so smaller frames are worth the probe too) */
            F3IterateParams Q = P;
            Q.probe_iters = 64;
            Q.sweep_on = 0;
            Q.warp_uniform = 0;
            Q.badvals = ctx->d_badvals + 4;
            int pgrid = ctx->sms < grid ? ctx->sms : grid;
            const uint64_t lim = (uint64_t)P.group_begin + (uint64_t)pgrid * (uint64_t)nranks;
            if (lim < Q.n_groups) Q.n_groups = (uint32_t)lim;
            /* the counters were zeroed by f3b_frame_begin / the previous probe's read-back */
            if (!f3b_launch_iterate(Q, consts.data(), j1 - j0, ctx->count_channel != 0, simple, true, pgrid,
                                    group_size, max_blob, ctx->stream))
               return fail(ctx, "probe launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            launches++;
            unsigned long long pc[2] = {0, 0};
            CK(cudaMemcpyAsync(pc, ctx->d_badvals + 2, sizeof(pc), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            aggregate = pc[0] > 0 && pc[1] * 8 >= pc[0];
            CK(cudaMemsetAsync(ctx->d_badvals + 2, 0, 3 * sizeof(unsigned long long), ctx->stream));
         }
      }
      const int nl = f3b_launch_iterate(P, consts.data(), j1 - j0, ctx->count_channel != 0, simple,
                                        aggregate, grid, group_size, smem_bytes, st);
      if (!nl) return fail(ctx, "chaos-game launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      launches += nl;
   }
   if (fork && !binned)
      for (int i = 0; i < F3B_SIDE_STREAMS; i++) {
         CK(cudaEventRecord(ctx->join_ev[i], ctx->side[i]));
         CK(cudaStreamWaitEvent(ctx->stream, ctx->join_ev[i], 0));
      }
   if (binned) { /* every tile has been merged already */
      ctx->staging_dirty = false;
      ctx->hot_measured = true;
      ctx->hot_from_merge = true;
      ctx->hot_total = 255.0 * (double)ctx->hot_iters / (double)nranks;
      return stage_end(ctx, launches);
   }
   /* staging copies -> exact buckets; the merge sees this rank's whole histogram (1/nranks of
    * the hits) and also records the fullest bucket */
   F3MergeParams M;
   M.color = (float4 *)ctx->hist_color.p;
   M.count = ctx->count_channel ? (float *)ctx->hist_count.p : nullptr;
   M.exact_color = ctx->exact_color.p;
   M.exact_count = ctx->count_channel ? ctx->exact_count.p : nullptr;
   M.f64 = ctx->f64;
   M.elems = ctx->hist_elems;
   M.lay = ctx->lay;
   M.hot_max = (unsigned int *)(ctx->d_badvals + 1);
   M.exact_flags = ctx->d_badvals + F3B_CTR_LOST;
   M.sat_list = (const uint32_t *)ctx->sat_list.p;
   if (P.sweep_on || aggregate) ctx->swept = true; /* integer atomics may have wrapped */
   M.sat_check = ctx->swept ? 1 : 0;
   const int nm = f3b_launch_hist_merge(M, ctx->stream);
   if (nm < 0) return fail(ctx, "histogram merge launch failed: %s", cudaGetErrorString(cudaGetLastError()));
   ctx->staging_dirty = false;
   ctx->hot_measured = true;
   ctx->hot_from_merge = true;
   ctx->hot_total = 255.0 * (double)ctx->hot_iters / (double)nranks;
   launches += nm;
   return stage_end(ctx, launches);
}

int f3b_orbits(f3b_ctx *ctx, const void *blob, size_t blob_bytes, int n_chains, int n_per_chain, int fuse,
               uint64_t seed, const float *start_xy, float *out_xycv, uint64_t *badvals) {
   if (!ctx || !blob || !out_xycv || n_chains <= 0 || n_per_chain <= 0 || fuse < 0)
      return fail(ctx, "f3b_orbits: bad arguments");
   CK(cudaSetDevice(ctx->device));
   const uint32_t *w = (const uint32_t *)blob;
   if (blob_bytes < F3B_H_SIZE * 4 || (size_t)w[F3B_H_WORDS] * 4 > blob_bytes || w[F3B_H_WORDS] > F3B_MAX_BLOB_WORDS)
      return fail(ctx, "f3b_orbits: malformed blob");
   const size_t nout = (size_t)n_chains * n_per_chain;
   const size_t bb = ((size_t)w[F3B_H_WORDS] * 4 + 15) & ~(size_t)15;
   const size_t sb = start_xy ? (((size_t)n_chains * 2 * sizeof(float) + 15) & ~(size_t)15) : 0;
   if (ensure(ctx, ctx->xfer, bb + sb + 16 + nout * sizeof(float4))) return 1;
   char *base = (char *)ctx->xfer.p;
   unsigned long long *d_bad = (unsigned long long *)(base + bb + sb);
   float4 *d_out = (float4 *)(base + bb + sb + 16);
   CK(cudaMemcpyAsync(base, blob, (size_t)w[F3B_H_WORDS] * 4, cudaMemcpyHostToDevice, ctx->stream));
   if (start_xy) CK(cudaMemcpyAsync(base + bb, start_xy, (size_t)n_chains * 2 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
   CK(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), ctx->stream));
   if (!f3b_launch_orbits((const uint32_t *)base, (size_t)w[F3B_H_WORDS] * 4, n_chains, n_per_chain, fuse, seed,
                          start_xy ? (const float *)(base + bb) : nullptr, d_out, d_bad, ctx->stream))
      return fail(ctx, "orbit launch failed: %s", cudaGetErrorString(cudaGetLastError()));
   CK(cudaMemcpyAsync(out_xycv, d_out, nout * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
   unsigned long long bad = 0;
   CK(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
   CK(cudaStreamSynchronize(ctx->stream));
   if (badvals) *badvals = bad;
   return 0;
}

int f3b_debug_apply_xform(f3b_ctx *ctx, const void *blob, size_t blob_bytes, int xform_index,
                          const float *in_xyc, int n, float *out_xycvb, uint64_t seed) {
   if (!ctx || !blob || !in_xyc || !out_xycvb || n <= 0) return fail(ctx, "debug_apply: bad arguments");
   CK(cudaSetDevice(ctx->device));
   const uint32_t *w = (const uint32_t *)blob;
   if (blob_bytes < F3B_H_SIZE * 4 || (size_t)w[F3B_H_WORDS] * 4 > blob_bytes ||
       w[F3B_H_WORDS] > F3B_MAX_BLOB_WORDS)
      return fail(ctx, "debug_apply: malformed blob");
   if (xform_index < 0 || (uint32_t)xform_index >= w[F3B_H_NXF])
      return fail(ctx, "debug_apply: xform index %d out of range", xform_index);
   void *d_blob = nullptr, *d_in = nullptr, *d_out = nullptr;
   cudaError_t e = cudaMalloc(&d_blob, blob_bytes);
   if (e == cudaSuccess) e = cudaMalloc(&d_in, sizeof(float) * 3 * (size_t)n);
   if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(float) * 5 * (size_t)n);
   if (e == cudaSuccess) e = cudaMemcpyAsync(d_blob, blob, blob_bytes, cudaMemcpyHostToDevice, ctx->stream);
   if (e == cudaSuccess)
      e = cudaMemcpyAsync(d_in, in_xyc, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
   if (e == cudaSuccess &&
       !f3b_launch_debug_apply((const uint32_t *)d_blob, (size_t)w[F3B_H_WORDS] * 4, xform_index,
                               (const float *)d_in, n, (float *)d_out, seed, ctx->stream))
      e = cudaGetLastError();
   if (e == cudaSuccess)
      e = cudaMemcpyAsync(out_xycvb, d_out, sizeof(float) * 5 * (size_t)n, cudaMemcpyDeviceToHost,
                          ctx->stream);
   if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
   cudaFree(d_blob);
   cudaFree(d_in);
   cudaFree(d_out);
   if (e != cudaSuccess) return fail(ctx, "debug_apply failed: %s", cudaGetErrorString(e));
   return 0;
}

int f3b_badvals(f3b_ctx *ctx, double *out) {
   unsigned long long v[F3B_CTR_TICKET] = {0};
   CK(cudaSetDevice(ctx->device));
   CK(cudaMemcpyAsync(v, ctx->d_badvals, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
   ctx->d2h_bytes += sizeof(v);
   CK(cudaStreamSynchronize(ctx->stream));
   *out = (double)v[0];
   ctx->lost_exact = v[F3B_CTR_LOST];
   ctx->saturated = v[F3B_CTR_LOST + 1];
   /* end of frame: remember how concentrated this genome's histogram was (see f3b_iterate) */
   if (ctx->hot_total > 0.0 && ctx->hot_measured) {
      ctx->hot_share = (double)(uint32_t)v[1] / ctx->hot_total;
      ctx->hot_share_signature = ctx->hot_signature;
   }
   return 0;
}

/* all device counters as of now (waits for the stream); for tools */
int f3b_debug_counters(f3b_ctx *ctx, uint64_t out[16]) {
   if (!ctx || !out) return 1;
   CK(cudaSetDevice(ctx->device));
   unsigned long long v[F3B_NCOUNTERS];
   CK(cudaMemcpyAsync(v, ctx->d_badvals, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
   CK(cudaStreamSynchronize(ctx->stream));
   for (int i = 0; i < F3B_NCOUNTERS; i++) out[i] = v[i];
   return 0;
}

int f3b_exactness(f3b_ctx *ctx, uint64_t *lost, uint64_t *saturated) {
   if (!ctx) return 1;
   if (lost) *lost = ctx->lost_exact;
   if (saturated) *saturated = ctx->saturated;
   return 0;
}

static double host_ms() {
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int f3b_density_rows(f3b_ctx *ctx, const f3b_density_desc *d, int row0, int row1) {
   if (!ctx || !d) return 1;
   const bool timing = getenv("F3B_TIMING") != nullptr;
   const double tt0 = host_ms();
   double tt1 = tt0, tt2 = tt0;
   if (!ctx->accum.p) return fail(ctx, "f3b_density: f3b_frame_begin not called");
   CK(cudaSetDevice(ctx->device));
   if (row0 < 0) row0 = 0;
   if (row1 > ctx->height) row1 = ctx->height;

   F3DensityParams P;
   memset(&P, 0, sizeof(P));
   P.hist.color = ctx->exact_color.p;
   P.hist.count = ctx->count_channel ? ctx->exact_count.p : nullptr;
   P.hist.f64 = ctx->f64;
   P.hist.pitch = ctx->lay.pitch;
   P.hist.row0 = 0;
   P.hist.rows = ctx->height;
   if (ctx->view_reduced) {
      P.hist.color = ctx->reduced_color.p;
      P.hist.count = ctx->count_channel ? ctx->reduced_count.p : nullptr;
      P.hist.row0 = ctx->view_row0;
      P.hist.rows = ctx->view_rows;
   }
   P.accum = (float4 *)ctx->accum.p;
   P.width = ctx->width;
   P.height = ctx->height;
   P.row_begin = row0;
   P.row_end = row1;
   P.oversample = d->oversample;
   P.k1 = (float)d->k1;
   P.k2 = (float)d->k2;
   P.int_accum = d->int_accum;
   P.n_kernels = 0;

   if (d->n_kernels > 0) {
      if (!d->coefs || !d->widths || d->kernel_size <= 0)
         return fail(ctx, "f3b_density: missing DE tables");
      if (d->max_filter_index >= d->n_kernels)
         return fail(ctx, "f3b_density: max_filter_index %d >= n_kernels %d", d->max_filter_index,
                     d->n_kernels);
      const size_t nc = (size_t)d->n_kernels * d->kernel_size;
      std::vector<float> coefs(nc);
      std::vector<uint8_t> radius(d->n_kernels);
      int maxr = 0;
      for (size_t i = 0; i < nc; i++) coefs[i] = (float)d->coefs[i];
      /* largest half-width the octant table can address: kernel_size = (h+1)(h+2)/2 */
      int half = 0;
      while ((half + 2) * (half + 3) / 2 <= d->kernel_size) half++;
      std::vector<float> kparam(2 * (size_t)d->n_kernels);
      std::vector<uint16_t> d2max(d->n_kernels);
      for (int k = 0; k < d->n_kernels; k++) {
         int r = (int)ceil(d->widths[k]) - 1; /* arr_filt_width, rect.c:123 */
         if (r < 0) r = 0;
         if (r > half) r = half;
         if (r > 254) return fail(ctx, "f3b_density: kernel radius %d too large", r);
         radius[k] = (uint8_t)r;
         if (k <= d->max_filter_index && r > maxr) maxr = r;
         /* closed form of the kernel for the staged gather (f3b_post.cu) */
         const double h = d->widths[k];
         kparam[2 * k] = h > 0.0 ? (float)(4.5 * 1.4426950408889634 / (h * h)) : 0.0f;
         kparam[2 * k + 1] = (float)d->coefs[(size_t)k * d->kernel_size];
         int m = 0;
         for (int j = 0; j <= r; j++)
            for (int i = 0; i <= j; i++)
               if (d->coefs[(size_t)k * d->kernel_size + (size_t)j * (j + 1) / 2 + i] != 0.0 && j * j + i * i > m)
                  m = j * j + i * i;
         d2max[k] = (uint16_t)(m > 65535 ? 65535 : m);
      }
      /* the tables only change with the estimator parameters: upload (and wait, the host
       * vectors are temporaries) only when they differ from what the device already holds */
      if (coefs != ctx->de_coefs_host || radius != ctx->de_radius_host) {
         if (ensure(ctx, ctx->de_coefs, nc * sizeof(float))) return 1;
         if (ensure(ctx, ctx->de_radius, (size_t)d->n_kernels)) return 1;
         if (ensure(ctx, ctx->de_kparam, (size_t)d->n_kernels * 2 * sizeof(float))) return 1;
         if (ensure(ctx, ctx->de_d2max, (size_t)d->n_kernels * sizeof(uint16_t))) return 1;
         CK(cudaMemcpyAsync(ctx->de_kparam.p, kparam.data(), kparam.size() * sizeof(float), cudaMemcpyHostToDevice,
                            ctx->stream));
         CK(cudaMemcpyAsync(ctx->de_d2max.p, d2max.data(), d2max.size() * sizeof(uint16_t), cudaMemcpyHostToDevice,
                            ctx->stream));
         CK(cudaMemcpyAsync(ctx->de_coefs.p, coefs.data(), nc * sizeof(float), cudaMemcpyHostToDevice,
                            ctx->stream));
         CK(cudaMemcpyAsync(ctx->de_radius.p, radius.data(), (size_t)d->n_kernels,
                            cudaMemcpyHostToDevice, ctx->stream));
         ctx->h2d_bytes += nc * sizeof(float) + (size_t)d->n_kernels;
         tt1 = host_ms();
         CK(cudaStreamSynchronize(ctx->stream));
         ctx->de_coefs_host.swap(coefs);
         ctx->de_radius_host.swap(radius);
      } else
         tt1 = host_ms();
      tt2 = host_ms();
      P.de_coefs = (const float *)ctx->de_coefs.p;
      P.de_radius = (const uint8_t *)ctx->de_radius.p;
      P.de_kparam = (const float2 *)ctx->de_kparam.p;
      P.de_d2max = (const uint16_t *)ctx->de_d2max.p;
      P.de_table_only = getenv("F3B_DE_TABLE") != nullptr;
      P.n_kernels = d->n_kernels;
      P.kernel_size = d->kernel_size;
      P.max_radius = maxr;
      P.max_filtered_counts = d->max_filtered_counts;
      P.max_filter_index = d->max_filter_index;
      P.curve = (float)d->curve;
   }
   if (ctx->view_reduced) {
      /* sources up to max_radius rows beyond the targets, each looking oversample/2 further */
      const int reach = P.n_kernels > 0 ? P.max_radius + d->oversample / 2 : 0;
      const int need0 = row0 - reach < 0 ? 0 : row0 - reach;
      const int need1 = row1 + reach > ctx->height ? ctx->height : row1 + reach;
      if (need0 < ctx->view_row0 || need1 > ctx->view_row0 + ctx->view_rows)
         return fail(ctx, "f3b_density_rows: rows [%d,%d) need buckets [%d,%d), the reduced strip holds [%d,%d)",
                     row0, row1, need0, need1, ctx->view_row0, ctx->view_row0 + ctx->view_rows);
   }
   if (P.n_kernels > 0) {
      const size_t tiles = (size_t)((ctx->width + 31) / 32) * ((ctx->height + 31) / 32);
      if (ensure(ctx, ctx->de_kidx, (size_t)ctx->width * ctx->height * sizeof(uint16_t))) return 1;
      if (ensure(ctx, ctx->de_tiles, tiles * sizeof(int))) return 1;
      P.de_kidx = (uint16_t *)ctx->de_kidx.p;
      P.tile_maxrad = (int *)ctx->de_tiles.p;
      /* whole-frame passes only (a strip sees a part of the histogram), and only when the
       * merge of the histogram copies has not measured it already */
      P.hot_max = (!ctx->hot_from_merge && row0 == 0 && row1 == ctx->height) ? (unsigned int *)(ctx->d_badvals + 1) : nullptr;
      if (P.hot_max) ctx->hot_measured = true;
   }
   if (ctx->accum_zero_pending && P.n_kernels > 0 && row0 == 0 && row1 == ctx->height) {
      P.overwrite_accum = 1;
      ctx->accum_zero_pending = false;
   } else if (accum_materialize(ctx)) {
      return 1;
   }
   if (stage_begin(ctx, F3B_STAGE_DENSITY)) return 1;
   const int nl = f3b_launch_density(P, ctx->stream);
   if (!nl) return fail(ctx, "density launch failed: %s", cudaGetErrorString(cudaGetLastError()));
   if (timing)
      fprintf(stderr, "f3b timing (density host ms): tables+copies %.3f  wait for the stream %.3f  launch %.3f\n",
              tt1 - tt0, tt2 - tt1, host_ms() - tt2);
   return stage_end(ctx, nl);
}

int f3b_density(f3b_ctx *ctx, const f3b_density_desc *d) {
   return f3b_density_rows(ctx, d, 0, ctx ? ctx->height : 0);
}

int f3b_finish_device(f3b_ctx *ctx, const f3b_final_desc *d, void **dev_out) {
   return f3b_finish_device_rows(ctx, d, 0, d ? d->image_height : 0, 0, dev_out);
}

static int finish_rows_impl(f3b_ctx *ctx, const f3b_final_desc *d, int row0, int row1, int alloc_rows,
                            void *dst_image, void **dev_out);

int f3b_finish_device_rows(f3b_ctx *ctx, const f3b_final_desc *d, int row0, int row1, int alloc_rows,
                           void **dev_out) {
   return finish_rows_impl(ctx, d, row0, row1, alloc_rows, nullptr, dev_out);
}

int f3b_finish_rows_into(f3b_ctx *ctx, const f3b_final_desc *d, int row0, int row1, void *dst_image) {
   if (!dst_image) return fail(ctx, "f3b_finish_rows_into: null image");
   return finish_rows_impl(ctx, d, row0, row1, 0, dst_image, nullptr);
}

int f3b_image_reserve(f3b_ctx *ctx, const f3b_final_desc *d, void **dev_image) {
   if (!ctx || !d || !dev_image) return 1;
   CK(cudaSetDevice(ctx->device));
   const size_t ibytes = (size_t)d->image_width * d->image_height * d->nchan * d->bytes_per_channel;
   if (ensure(ctx, ctx->image, ibytes)) return 1;
   *dev_image = ctx->image.p;
   return 0;
}

static int finish_rows_impl(f3b_ctx *ctx, const f3b_final_desc *d, int row0, int row1, int alloc_rows,
                            void *dst_image, void **dev_out) {
   if (!ctx || !d) return 1;
   if (!ctx->accum.p) return fail(ctx, "f3b_finish: f3b_frame_begin not called");
   CK(cudaSetDevice(ctx->device));
   if (accum_materialize(ctx)) return 1;
   if (d->nchan != 3 && d->nchan != 4) return fail(ctx, "f3b_finish: nchan must be 3 or 4");
   if (d->bytes_per_channel != 1 && d->bytes_per_channel != 2)
      return fail(ctx, "f3b_finish: bytes_per_channel must be 1 or 2");
   if (d->filter_width <= 0 || !d->filter) return fail(ctx, "f3b_finish: missing spatial filter");
   /* the filter footprint of the last pixel must stay inside the bucket grid */
   const long xmax = (long)d->de_offset + (long)(d->image_width - 1) * d->oversample + d->filter_width;
   const long ymax = (long)d->de_offset + (long)(d->image_height - 1) * d->oversample + d->filter_width;
   if (xmax > ctx->width || ymax > ctx->height)
      return fail(ctx, "f3b_finish: filter footprint %ldx%ld exceeds buffer %dx%d", xmax, ymax,
                  ctx->width, ctx->height);
   const size_t fbytes = sizeof(double) * (size_t)d->filter_width * d->filter_width;
   if (row0 < 0) row0 = 0;
   if (row1 > d->image_height) row1 = d->image_height;
   if (alloc_rows < d->image_height) alloc_rows = d->image_height;
   const size_t ibytes = (size_t)d->image_width * alloc_rows * d->nchan * d->bytes_per_channel;
   if (!dst_image && ensure(ctx, ctx->image, ibytes)) return 1;
   {
      const size_t nf = (size_t)d->filter_width * d->filter_width;
      if (ctx->filter_host.size() != nf || memcmp(ctx->filter_host.data(), d->filter, fbytes) != 0) {
         if (ensure(ctx, ctx->filter, fbytes)) return 1;
         ctx->filter_host.assign(d->filter, d->filter + nf);
         /* from the context's own copy, which outlives the call */
         CK(cudaMemcpyAsync(ctx->filter.p, ctx->filter_host.data(), fbytes, cudaMemcpyHostToDevice,
                            ctx->stream));
         ctx->h2d_bytes += fbytes;
         CK(cudaStreamSynchronize(ctx->stream));
      }
   }

   F3FinalParams P;
   memset(&P, 0, sizeof(P));
   P.accum = (float4 *)ctx->accum.p;
   P.width = ctx->width;
   P.height = ctx->height;
   P.image_width = d->image_width;
   P.image_height = d->image_height;
   P.oversample = d->oversample;
   P.de_offset = d->de_offset;
   P.filter = (const double *)ctx->filter.p;
   P.filter_width = d->filter_width;
   P.earlyclip = d->earlyclip;
   P.nchan = d->nchan;
   P.transp = d->transp;
   P.bytes_per_channel = d->bytes_per_channel;
   P.g = 1.0 / d->gamma; /* rect.c:1071 */
   P.linrange = d->gam_lin_thresh;
   P.vibrancy = d->vibrancy;
   P.highpow = d->highlight_power;
   for (int i = 0; i < 3; i++) P.background[i] = d->background[i];
   P.out = dst_image ? dst_image : ctx->image.p;
   /* 8-bit images of up to 65535 x 65535 pixels: float pass + double fix-up of the listed
    * pixels (F3B_FINAL_DOUBLE=1: everything in double, for A/B) */
   P.fixup_list = nullptr;
   P.fixup_count = nullptr;
   if (d->bytes_per_channel == 1 && d->image_width < 65536 && d->image_height < 65536 &&
       !getenv("F3B_FINAL_DOUBLE")) {
      const size_t rows = (size_t)(row1 > row0 ? row1 - row0 : 0);
      if (ensure(ctx, ctx->fixup, (rows * d->image_width + 4) * sizeof(uint32_t))) return 1;
      P.fixup_count = (unsigned int *)ctx->fixup.p;
      P.fixup_list = (uint32_t *)ctx->fixup.p + 4;
      CK(cudaMemsetAsync(ctx->fixup.p, 0, 4 * sizeof(uint32_t), ctx->stream));
   }
   P.out_row_begin = row0;
   P.out_row_end = row1;

   if (stage_begin(ctx, F3B_STAGE_FINAL)) return 1;
   const int nl = f3b_launch_final(P, ctx->stream);
   if (!nl) return fail(ctx, "final-pass launch failed: %s", cudaGetErrorString(cudaGetLastError()));
   if (stage_end(ctx, nl)) return 1;
   if (dev_out) *dev_out = ctx->image.p;
   return 0;
}

int f3b_finish(f3b_ctx *ctx, const f3b_final_desc *d, void *out, size_t row_stride) {
   if (!out) return fail(ctx, "f3b_finish: null output");
   if (f3b_finish_device(ctx, d, nullptr)) return 1;
   return f3b_image_download(ctx, d, out, row_stride);
}

int f3b_image_download(f3b_ctx *ctx, const f3b_final_desc *d, void *out, size_t row_stride) {
   if (!ctx || !d || !out) return fail(ctx, "f3b_image_download: null argument");
   CK(cudaSetDevice(ctx->device));
   void *dev = ctx->image.p;
   if (!dev) return fail(ctx, "f3b_image_download: no image (f3b_finish_device not called)");
   const size_t row = (size_t)d->image_width * d->nchan * d->bytes_per_channel;
   if (row_stride < row) return fail(ctx, "f3b_finish: row stride too small");
   CK(cudaMemcpy2DAsync(out, row_stride, dev, row, row, (size_t)d->image_height,
                        cudaMemcpyDeviceToHost, ctx->stream));
   ctx->d2h_bytes += row * (size_t)d->image_height;
   CK(cudaStreamSynchronize(ctx->stream));
   return 0;
}

int f3b_hist_upload_u32(f3b_ctx *ctx, const uint32_t *buckets5) {
   if (!ctx || !ctx->exact_color.p) return 1;
   CK(cudaSetDevice(ctx->device));
   const size_t n = (size_t)ctx->width * ctx->height;
   if (ensure(ctx, ctx->xfer, n * 5 * sizeof(uint32_t))) return 1;
   CK(cudaMemcpyAsync(ctx->xfer.p, buckets5, n * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
   ctx->h2d_bytes += n * 5 * sizeof(uint32_t);
   if (!f3b_launch_hist_from_u32((const uint32_t *)ctx->xfer.p, ctx->exact_color.p,
                                 ctx->count_channel ? ctx->exact_count.p : nullptr, ctx->f64, ctx->width,
                                 ctx->height, ctx->lay.pitch, ctx->stream))
      return fail(ctx, "hist upload failed: %s", cudaGetErrorString(cudaGetLastError()));
   CK(cudaStreamSynchronize(ctx->stream)); /* the caller's buffer is free again */
   return 0;
}

int f3b_hist_download_u32(f3b_ctx *ctx, uint32_t *buckets5) {
   if (!ctx || !ctx->exact_color.p) return 1;
   CK(cudaSetDevice(ctx->device));
   const size_t n = (size_t)ctx->width * ctx->height;
   if (ensure(ctx, ctx->xfer, n * 5 * sizeof(uint32_t))) return 1;
   if (!f3b_launch_hist_to_u32(ctx->exact_color.p, ctx->count_channel ? ctx->exact_count.p : nullptr, ctx->f64,
                               (uint32_t *)ctx->xfer.p, ctx->width, ctx->height, ctx->lay.pitch, ctx->stream))
      return fail(ctx, "hist download failed: %s", cudaGetErrorString(cudaGetLastError()));
   CK(cudaMemcpyAsync(buckets5, ctx->xfer.p, n * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
   ctx->d2h_bytes += n * 5 * sizeof(uint32_t);
   CK(cudaStreamSynchronize(ctx->stream));
   return 0;
}

int f3b_accum_upload(f3b_ctx *ctx, const float *acc4) {
   if (!ctx || !ctx->accum.p) return 1;
   CK(cudaSetDevice(ctx->device));
   const size_t n = (size_t)ctx->width * ctx->height;
   ctx->accum_zero_pending = false;
   CK(cudaMemcpyAsync(ctx->accum.p, acc4, n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
   CK(cudaStreamSynchronize(ctx->stream));
   return 0;
}

int f3b_accum_download(f3b_ctx *ctx, float *acc4) {
   if (!ctx || !ctx->accum.p) return 1;
   CK(cudaSetDevice(ctx->device));
   const size_t n = (size_t)ctx->width * ctx->height;
   if (accum_materialize(ctx)) return 1;
   CK(cudaMemcpyAsync(acc4, ctx->accum.p, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
   CK(cudaStreamSynchronize(ctx->stream));
   return 0;
}

void *f3b_hist_color_ptr(f3b_ctx *ctx) { return ctx ? ctx->exact_color.p : nullptr; }
size_t f3b_hist_elems(f3b_ctx *ctx) { return ctx ? ctx->hist_elems : 0; }
void *f3b_hist_count_ptr(f3b_ctx *ctx) {
   return (ctx && ctx->count_channel) ? ctx->exact_count.p : nullptr;
}
void *f3b_accum_ptr(f3b_ctx *ctx) {
   if (!ctx || accum_materialize(ctx)) return nullptr;
   return ctx->accum.p;
}
void *f3b_stream_ptr(f3b_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

/* ---- several GPUs on one frame ------------------------------------------------------ */
int f3b_peer_enable(f3b_ctx *ctx, f3b_ctx *peer) {
   if (!ctx || !peer) return 1;
   if (ctx->device == peer->device) return 0;
   CK(cudaSetDevice(ctx->device));
   int can = 0;
   CK(cudaDeviceCanAccessPeer(&can, ctx->device, peer->device));
   if (!can) return fail(ctx, "device %d cannot access device %d directly", ctx->device, peer->device);
   cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
   if (e == cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
      e = cudaSuccess;
   }
   if (e != cudaSuccess) return fail(ctx, "cudaDeviceEnablePeerAccess(%d) failed: %s", peer->device, cudaGetErrorString(e));
   return 0;
}

int f3b_hist_reduce_rows(f3b_ctx *ctx, const f3b_peer_hist *peers, int npeers, int row0, int row1) {
   if (!ctx || !peers || npeers < 1 || npeers > F3B_MAX_RANKS) return fail(ctx, "f3b_hist_reduce_rows: bad arguments");
   if (!ctx->exact_color.p) return fail(ctx, "f3b_hist_reduce_rows: f3b_frame_begin not called");
   CK(cudaSetDevice(ctx->device));
   if (row0 < 0) row0 = 0;
   if (row1 > ctx->height) row1 = ctx->height;
   if (row1 <= row0) return fail(ctx, "f3b_hist_reduce_rows: empty row range");
   F3ReduceParams R;
   memset(&R, 0, sizeof(R));
   R.npeers = npeers;
   R.f64 = ctx->f64;
   for (int p = 0; p < npeers; p++) {
      if (!peers[p].color || (ctx->count_channel && !peers[p].count))
         return fail(ctx, "f3b_hist_reduce_rows: peer %d has no buckets", p);
      R.color[p] = peers[p].color;
      R.count[p] = ctx->count_channel ? peers[p].count : nullptr;
   }
   R.first = (size_t)row0 * ctx->lay.pitch;
   R.n = (size_t)(row1 - row0) * ctx->lay.pitch;
   if (ensure(ctx, ctx->reduced_color, R.n * (ctx->f64 ? sizeof(double4) : sizeof(uint4)))) return 1;
   if (ctx->count_channel && ensure(ctx, ctx->reduced_count, R.n * (ctx->f64 ? sizeof(double) : sizeof(uint32_t)))) return 1;
   R.out_color = ctx->reduced_color.p;
   R.out_count = ctx->count_channel ? ctx->reduced_count.p : nullptr;
   if (stage_begin(ctx, F3B_STAGE_EXCHANGE)) return 1;
   if (!f3b_launch_hist_reduce(R, ctx->stream))
      return fail(ctx, "strip reduction launch failed: %s", cudaGetErrorString(cudaGetLastError()));
   if (stage_end(ctx, 1)) return 1;
   ctx->view_reduced = true;
   ctx->view_row0 = row0;
   ctx->view_rows = row1 - row0;
   ctx->peer_bytes += (uint64_t)R.n * (uint64_t)(npeers - 1) * (ctx->f64 ? 2 : 1) * (ctx->count_channel ? 20 : 16);
   return 0;
}

int f3b_group_barrier(f3b_ctx *const *ctxs, int n) {
   if (!ctxs || n < 1) return 1;
   for (int i = 0; i < n; i++) {
      f3b_ctx *ctx = ctxs[i];
      if (!ctx) return 1;
      CK(cudaSetDevice(ctx->device));
      if (!ctx->bar_ev) CK(cudaEventCreateWithFlags(&ctx->bar_ev, cudaEventDisableTiming));
      CK(cudaEventRecord(ctx->bar_ev, ctx->stream));
   }
   for (int i = 0; i < n; i++) {
      f3b_ctx *ctx = ctxs[i];
      CK(cudaSetDevice(ctx->device));
      for (int j = 0; j < n; j++)
         if (j != i) CK(cudaStreamWaitEvent(ctx->stream, ctxs[j]->bar_ev, 0));
   }
   return 0;
}

int f3b_ipc_export(f3b_ctx *ctx, const void *dev_ptr, unsigned char handle[64]) {
   if (!ctx || !dev_ptr || !handle) return 1;
   CK(cudaSetDevice(ctx->device));
   static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
   cudaIpcMemHandle_t h;
   CK(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
   memcpy(handle, &h, 64);
   return 0;
}
int f3b_ipc_open(f3b_ctx *ctx, const unsigned char handle[64], void **dev_ptr) {
   if (!ctx || !handle || !dev_ptr) return 1;
   CK(cudaSetDevice(ctx->device));
   cudaIpcMemHandle_t h;
   memcpy(&h, handle, 64);
   CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
   return 0;
}
int f3b_ipc_close(f3b_ctx *ctx, void *dev_ptr) {
   if (!ctx || !dev_ptr) return 1;
   CK(cudaSetDevice(ctx->device));
   CK(cudaIpcCloseMemHandle(dev_ptr));
   return 0;
}
void *f3b_image_ptr(f3b_ctx *ctx) { return ctx ? ctx->image.p : nullptr; }
int f3b_device_index(f3b_ctx *ctx) { return ctx ? ctx->device : -1; }

int f3b_stage_stats(f3b_ctx *ctx, int stage, double *ms, uint64_t *launches) {
   if (!ctx || stage < 0 || stage >= F3B_NSTAGES) return 1;
   CK(cudaSetDevice(ctx->device));
   if (stats_collect(ctx)) return 1;
   if (ms) *ms = ctx->stage_ms[stage];
   if (launches) *launches = ctx->stage_launches[stage];
   return 0;
}

int f3b_xfer_stats(f3b_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
   if (!ctx) return 1;
   if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
   if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
   return 0;
}

int f3b_timer_start(f3b_ctx *ctx) {
   if (!ctx) return 1;
   CK(cudaSetDevice(ctx->device));
   if (!ctx->timer0) {
      CK(cudaEventCreate(&ctx->timer0));
      CK(cudaEventCreate(&ctx->timer1));
   }
   CK(cudaEventRecord(ctx->timer0, ctx->stream));
   return 0;
}

int f3b_timer_stop(f3b_ctx *ctx, double *ms) {
   if (!ctx || !ctx->timer0) return 1;
   CK(cudaSetDevice(ctx->device));
   CK(cudaEventRecord(ctx->timer1, ctx->stream));
   CK(cudaEventSynchronize(ctx->timer1));
   float f = 0.f;
   CK(cudaEventElapsedTime(&f, ctx->timer0, ctx->timer1));
   if (ms) *ms = (double)f;
   return 0;
}

void f3b_stats_reset(f3b_ctx *ctx) {
   if (!ctx) return;
   cudaSetDevice(ctx->device);
   stats_collect(ctx);
   ctx->h2d_bytes = ctx->d2h_bytes = ctx->peer_bytes = 0;
   for (int i = 0; i < F3B_NSTAGES; i++) {
      ctx->stage_ms[i] = 0.0;
      ctx->stage_launches[i] = 0;
   }
}

} /* extern "C" */
