/*
 * flam3_b200.h — C-ABI of libflam3_b200.so, the B200 (sm_100a) implementation of flam3's
 * render hot path. Plain pointers and sizes only; no CUDA or torch types.
 *
 * Two levels:
 *
 *  (1) The drop-in symbols `flam3_render` / `flam3_render_memory_required` with the
 *      reference's exact ABI are declared in flam3_b200_dropin.h (they need the
 *      reference's public header flam3.h for flam3_frame) and implemented in
 *      flam3_b200/csrc/f3b_render.c on top of this file.
 *
 *  (2) This file: the stage-level interface that replaces rect.c's
 *      render_rectangle / iter_thread / de_thread machinery. One f3b_ctx owns one GPU's
 *      buffers: the histogram ("buckets", rect.c:699), the accumulator (rect.c:700) and
 *      the output image. A frame is rendered as
 *
 *         f3b_frame_begin                      rect.c:685-710   allocate + clear
 *         for each batch:
 *            f3b_hist_clear                    rect.c:721
 *            f3b_iterate  (all temporal samples of the batch in one launch)
 *                                              rect.c:754-931 + iter_thread + flam3_iterate
 *            f3b_density                       rect.c:933-1058  k1/k2 log scale + DE
 *         f3b_finish                           rect.c:1067-1247 filter + tone map -> out
 *
 *      Multi-GPU callers shard f3b_iterate with (rank, nranks), sum the histograms
 *      through f3b_hist_*_ptr with their own collective, and run f3b_density/f3b_finish
 *      per strip.
 *
 * Every function returns 0 on success and non-zero on error (f3b_last_error has the
 * text) unless documented otherwise, mirroring the reference's 0/1 convention
 * (rect.c:598-606).
 */
#ifndef FLAM3_B200_H
#define FLAM3_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F3B_ABI_VERSION 3

typedef struct f3b_ctx f3b_ctx;

/* ---- context ------------------------------------------------------------------- */
int f3b_abi_version(void);
int f3b_ctx_create(int device, f3b_ctx **out);
void f3b_ctx_destroy(f3b_ctx *ctx);
const char *f3b_last_error(const f3b_ctx *ctx);
/* info[0]=SM count, info[1]=resident chaos-game CTAs per SM, info[2]=threads per CTA */
int f3b_device_info(f3b_ctx *ctx, int *info);
int f3b_sync(f3b_ctx *ctx);

/* ---- frame buffers (rect.c:685-710) ---------------------------------------------- */
/* width/height are the BUCKET grid (oversample*image + 2*gutter). count_channel=0 lets
 * the alpha channel double as the hit-count channel (valid when every palette alpha is 1 and
 * every temporal-filter weight is 1; f3b_render.c decides). `bits` is flam3_frame.bits and
 * selects the bucket type like flam3.c:3894-3932 does: 32 / 33 -> uint32[5] per bucket, adds
 * truncated, saturating at UINT_MAX (flam3.c:3772, 3824, 3838-3854); 64 -> double[5]
 * (flam3.c:3724). Clears buckets and accumulators. */
int f3b_frame_begin(f3b_ctx *ctx, int width, int height, int count_channel, int bits);
int f3b_hist_clear(f3b_ctx *ctx);
int f3b_accum_clear(f3b_ctx *ctx);

/* ---- chaos game + histogram (iter_thread rect.c:253-527, flam3_iterate flam3.c:232) - */
/* One entry per temporal sample; layout documented in flam3_b200/csrc/f3b_pack.h. */
typedef struct {
   uint64_t blob_offset; /* byte offset of the packed sample inside `blobs` (16-aligned) */
   uint64_t n_iters;     /* stored iterations (fic.batch_size * nthreads, rect.c:833-840) */
   uint32_t n_chains;    /* independent chains; 0 = let the library plan (f3b_plan_chains) */
   uint32_t fuse;        /* discarded iterations per chain (rect.c:266) */
   uint32_t first_group; /* filled by the library */
   uint32_t sample_id;   /* RNG stream id of this temporal sample */
} f3b_job;

/* Number of chains the library uses for n_iters stored iterations when the launch holds
 * n_jobs temporal samples, no chain may store more than max_chain_len iterations
 * (= flam3_frame.sub_batch_size, rect.c:279) and the caller asked for `workers` threads
 * (= flam3_frame.nthreads: the reference runs workers * ceil(n / workers / sub_batch_size)
 * sub-batches, rect.c:840 + 279-284; that is the minimum, see f3b_api.cu for the rounding). */
uint32_t f3b_plan_chains(f3b_ctx *ctx, uint64_t n_iters, uint32_t max_chain_len, uint32_t n_jobs,
                         uint32_t workers);

/* Runs the jobs' chains with chain-group index g where g % nranks == rank and adds the
 * hits into this context's buckets. floor_adds=1 truncates every add like the integer buckets
 * of bits 32/33 (flam3.c:3838-3854). seed selects the random streams. Jobs with n_chains == 0
 * are planned here, like f3b_plan_chains but rounding to the threads the kernel class that
 * runs the batch keeps resident (3 CTAs per SM for the classic-flame kernels, 2 for the
 * generic ones); `workers` = flam3_frame.nthreads. */
int f3b_iterate(f3b_ctx *ctx, const void *blobs, size_t blob_bytes, f3b_job *jobs, int n_jobs,
                uint64_t seed, int floor_adds, uint32_t max_chain_len, uint32_t workers, int rank,
                int nranks);
/* One application of xform `xform_index` of a packed sample to n caller-supplied points
 * (apply_xform, variations.c:2129-2399): in = {x, y, color} per point, out = {x, y, color,
 * vis_adjusted, bad} per point. For testing the variation maths in isolation. */
/* `blobs` may be page-locked host memory from f3b_pinned_alloc; the upload is then
 * asynchronous and f3b_iterate returns while it may still be reading `blobs`:
 * f3b_inputs_consumed waits until the buffer may be rewritten. (Pageable buffers are
 * consumed before f3b_iterate returns.) */
void *f3b_pinned_alloc(size_t bytes);
void f3b_pinned_free(void *p);
int f3b_inputs_consumed(f3b_ctx *ctx);

/* flam3_iterate (flam3.h:555, flam3.c:232-299) for n_chains independent chains at once: each
 * starts at start_xy[2 t .. 2 t + 1] (NULL: a random point of [-1,1]^2) with colour 0, discards
 * `fuse` iterations and stores n_per_chain samples. out_xycv (host) receives
 * {x, y, colour, vis_adjusted} of sample i of chain t at [(i * n_chains + t) * 4]; *badvals the
 * number of bad values met. `blob` is ONE packed sample (f3b_pack.h). What the analysis helpers
 * either side of the render need (flam3_estimate_bounding_box flam3.c:3574, flam3_colorhist
 * flam3.c:351; both provided by the drop-in on top of this). */
int f3b_orbits(f3b_ctx *ctx, const void *blob, size_t blob_bytes, int n_chains, int n_per_chain, int fuse,
               uint64_t seed, const float *start_xy, float *out_xycv, uint64_t *badvals);
int f3b_debug_apply_xform(f3b_ctx *ctx, const void *blob, size_t blob_bytes, int xform_index,
                          const float *in_xyc, int n, float *out_xycvb, uint64_t seed);
/* bad values since f3b_frame_begin (stat_struct.badvals, rect.c:1251); waits for the stream */
int f3b_badvals(f3b_ctx *ctx, double *out);
/* Integer-bucket bookkeeping of the frame, as of the last f3b_badvals: *lost = staged float
 * sums that were found at >= 2^24 before they could be moved into the uint32 buckets (0 = the
 * histogram is exactly the integer histogram the reference's adds would give); *saturated =
 * uint32 sums that passed UINT_MAX and were pinned there (flam3.c:3838-3854). */
int f3b_exactness(f3b_ctx *ctx, uint64_t *lost, uint64_t *saturated);

/* ---- log scale + density estimation (rect.c:933-1058, de_thread rect.c:57-251) ---- */
typedef struct {
   double k1, k2;              /* rect.c:933-937 */
   int oversample;
   int int_accum;              /* bits=32: integer accumulators (flam3.c:3772-3786) */
   /* flam3_de_helper (filters.c:272-407). n_kernels=0 disables DE (rect.c:947). */
   int n_kernels;              /* de_max_ind rows present in coefs/widths */
   int kernel_size;            /* coefficients per kernel, octant packed */
   int max_filtered_counts;
   int max_filter_index;
   double curve;               /* estimator_curve (rect.c:997) */
   const double *coefs;        /* host, n_kernels*kernel_size */
   const double *widths;       /* host, n_kernels */
} f3b_density_desc;
int f3b_density(f3b_ctx *ctx, const f3b_density_desc *d);
/* Same, restricted to bucket rows [row0,row1) of the accumulator (multi-GPU strips). */
int f3b_density_rows(f3b_ctx *ctx, const f3b_density_desc *d, int row0, int row1);

/* ---- spatial filter + tone map + quantise (rect.c:1067-1247) ---------------------- */
typedef struct {
   int image_width, image_height; /* output pixels (field-adjusted height) */
   int oversample, de_offset;
   int filter_width;
   const double *filter;          /* host, filter_width^2 (filters.c:217-270) */
   int earlyclip, nchan, transp, bytes_per_channel;
   double gamma;                  /* mean gamma over temporal samples (rect.c:1071) */
   double vibrancy;               /* mean vibrancy */
   double background[3];          /* mean background * 256 (rect.c:1078-1080) */
   double highlight_power;
   double gam_lin_thresh;
} f3b_final_desc;
/* Writes image rows to host memory `out`; consecutive rows are row_stride bytes apart
 * (2x the row size for interlaced fields, rect.c:616-624). */
int f3b_finish(f3b_ctx *ctx, const f3b_final_desc *d, void *out, size_t row_stride);
/* Same but leaves the image in device memory (tightly packed); *dev_out receives it. */
int f3b_finish_device(f3b_ctx *ctx, const f3b_final_desc *d, void **dev_out);
/* Multi-GPU strips: produce only image rows [row0,row1) (the rows of rect.c:1139's loop) in
 * a device image of at least alloc_rows rows; the other rows are left untouched. */
int f3b_finish_device_rows(f3b_ctx *ctx, const f3b_final_desc *d, int row0, int row1,
                           int alloc_rows, void **dev_out);
/* Copies the device image of the last f3b_finish_device[_rows] to host rows row_stride apart. */
int f3b_image_download(f3b_ctx *ctx, const f3b_final_desc *d, void *out, size_t row_stride);

/* ---- raw buffer access: tests, histogram exchange between GPUs --------------------- */
/* reference bits=33 bucket layout: uint32[5] = r,g,b,a,count per bucket (flam3.c:3824) */
int f3b_hist_upload_u32(f3b_ctx *ctx, const uint32_t *buckets5);
int f3b_hist_download_u32(f3b_ctx *ctx, uint32_t *buckets5);
/* accumulator: float[4] per bucket (abucket_float, flam3.c:3825) */
int f3b_accum_upload(f3b_ctx *ctx, const float *acc4);
int f3b_accum_download(f3b_ctx *ctx, float *acc4);
void *f3b_hist_color_ptr(f3b_ctx *ctx); /* device uint32[4][f3b_hist_elems] (double[4] for bits 64) */
void *f3b_hist_count_ptr(f3b_ctx *ctx); /* device uint32[f3b_hist_elems] (double) or NULL */
/* buckets in a histogram plane = height * row pitch (the pitch may exceed the width; padding
 * buckets stay zero): the number of elements an exchange between GPUs must cover. Use
 * f3b_hist_download_u32 / f3b_hist_upload_u32 for the reference's [height][width] layout. */
size_t f3b_hist_elems(f3b_ctx *ctx);
void *f3b_accum_ptr(f3b_ctx *ctx);      /* device float4[height*width] */
void *f3b_stream_ptr(f3b_ctx *ctx);     /* cudaStream_t the context launches on */

/* ---- several GPUs on one frame (rect.c:840 splits the sample budget over threads the
 * same way) ---------------------------------------------------------
 * Every rank runs its share of the chains into its own buckets (f3b_iterate's rank/nranks).
 * Then rank r owns the image rows of strip r: it sums, straight out of the peers' memory
 * over NVLink, all ranks' buckets of the bucket rows that strip needs (its rows plus the
 * halo the widest density-estimation kernel and the spatial filter reach: overlapping rows
 * are simply fetched by both neighbours), saturating like the buckets do, and the post-pass
 * of those rows reads the sum: a reduce-scatter by image strip fused into one kernel. Its
 * image rows go straight into the image of the rank that owns `out` (f3b_finish_rows_into
 * with a peer pointer). The caller orders the steps: all ranks' f3b_iterate before any
 * f3b_hist_reduce_rows, every reduce before any rank clears its buckets again
 * (f3b_group_barrier for contexts of one process; ranks in separate processes exchange
 * pointers with f3b_ipc_* and synchronise their streams themselves). */
typedef struct {
   void *color; /* f3b_hist_color_ptr of that rank, valid on this context's device */
   void *count; /* f3b_hist_count_ptr, NULL without a count plane */
} f3b_peer_hist;
/* lets this context's device read and write `peer`'s memory (same process) */
int f3b_peer_enable(f3b_ctx *ctx, f3b_ctx *peer);
/* buckets of rows [row0,row1) summed over peers[0..npeers) (this rank included) become what
 * f3b_density_rows / f3b_density read until the next f3b_hist_clear / f3b_frame_begin */
int f3b_hist_reduce_rows(f3b_ctx *ctx, const f3b_peer_hist *peers, int npeers, int row0, int row1);
/* every context's stream waits until all n streams have reached this point (one process) */
int f3b_group_barrier(f3b_ctx *const *ctxs, int n);
/* device image of d->image_height rows, allocated now so that peers can be told where it is */
int f3b_image_reserve(f3b_ctx *ctx, const f3b_final_desc *d, void **dev_image);
void *f3b_image_ptr(f3b_ctx *ctx);
/* like f3b_finish_device_rows, but the rows go into `dev_image` (own or a peer's) */
int f3b_finish_rows_into(f3b_ctx *ctx, const f3b_final_desc *d, int row0, int row1, void *dev_image);
/* CUDA IPC for ranks in separate processes: 64-byte handles of device allocations (the
 * pointers returned by f3b_hist_color_ptr / f3b_hist_count_ptr / f3b_image_ptr) */
int f3b_ipc_export(f3b_ctx *ctx, const void *dev_ptr, unsigned char handle[64]);
int f3b_ipc_open(f3b_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
int f3b_ipc_close(f3b_ctx *ctx, void *dev_ptr);
int f3b_device_index(f3b_ctx *ctx);

/* ---- instrumentation ------------------------------------------------------------- */
enum { F3B_STAGE_ITERATE = 0, F3B_STAGE_DENSITY = 1, F3B_STAGE_FINAL = 2, F3B_STAGE_EXCHANGE = 3,
       F3B_NSTAGES = 4 };
/* GPU milliseconds (CUDA events on the context stream) and launch count of a stage since
 * f3b_stats_reset; synchronises the stream. */
int f3b_stage_stats(f3b_ctx *ctx, int stage, double *ms, uint64_t *launches);
void f3b_stats_reset(f3b_ctx *ctx);
/* Bytes copied host->device / device->host by this context since f3b_stats_reset. */
int f3b_xfer_stats(f3b_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);
/* The context's 16 device counters as of now (waits for the stream): [0] bad values, [1] fullest
 * bucket's count, [6] staged sums seen at >= 2^24, [7] saturated sums, [10..12] where the former
 * were seen (cold sweep, hot list, merge). For tools. */
int f3b_debug_counters(f3b_ctx *ctx, uint64_t out[16]);
/* The layout of the float staging copies in front of the exact buckets for a width x height
 * grid: out = {entries per copy, copies, scramble mask m, multiplier a, its inverse}. Bucket
 * b = x + width * y sits at entry (b & ~m) | ((b * a) & m) of a copy (DESIGN.md section 3:
 * neighbouring buckets must not share L2 sectors). Needs no GPU. */
int f3b_debug_stage_layout(int width, int height, int count_channel, uint64_t out[5]);
/* A CUDA-event stopwatch on the context's stream: start records an event, stop records a
 * second one, waits for it and returns the elapsed GPU milliseconds. */
int f3b_timer_start(f3b_ctx *ctx);
int f3b_timer_stop(f3b_ctx *ctx, double *ms);

#ifdef __cplusplus
}
#endif
#endif
