/*
 * f3b_internal.h — launch wrappers shared between the CUDA translation units of
 * libflam3_b200.so. Not part of the public C-ABI (that is include/flam3_b200.h).
 */
#ifndef F3B_INTERNAL_H
#define F3B_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flam3_b200.h"
#include "f3b_pack.h"

typedef f3b_job f3b_sample_job;

/* Where bucket (x, y) lives in the histogram planes: rows are `pitch` buckets apart. While
 * the chaos game runs there are `replicas` copies of every plane, `replica_stride` buckets
 * apart, and a CTA adds into copy blockIdx.x % replicas; they are summed into copy 0 before
 * anything reads the histogram (f3b_api.cu: hist_layout_for says why). */
struct F3HistLayout {
   int pitch;
   uint32_t replicas;
   size_t replica_stride;
   /* Where bucket b (= x + pitch * y) sits inside a STAGING copy: its low bits are multiplied by
    * an odd constant, a bijection of every aligned block of smask + 1 entries (replica_stride is
    * a multiple of it). An attractor's hits cluster on neighbouring buckets; in the plain layout
    * those are the same 32-byte sectors and 128-byte lines, whose reductions L2 serialises. The
    * scrambled layout puts neighbours 2.8 KB apart: C2's chaos game 24.0 -> 18.2 ms for the same
    * instruction stream (profiles/r02_staging_scramble_ab.txt). smask = 0: plain layout. The exact
    * buckets, and everything after the chaos game, stay in the plain layout. */
   uint32_t smask, smul, sinv;
};
#ifdef __CUDACC__
__device__ __forceinline__ size_t f3b_hist_index(const F3HistLayout &L, int x, int y) {
   return (size_t)x + (size_t)L.pitch * (size_t)y;
}
/* bucket -> entry of a staging copy, and back */
__device__ __forceinline__ uint32_t f3b_stage_index(const F3HistLayout &L, uint32_t b) {
   return (b & ~L.smask) | ((b * L.smul) & L.smask);
}
__device__ __forceinline__ uint32_t f3b_stage_bucket(const F3HistLayout &L, uint32_t s) {
   return (s & ~L.smask) | ((s * L.sinv) & L.smask);
}
#endif

/* The buckets proper (what the post-pass, the exchange between GPUs and the tests see) are
 * EXACT: uint32[4] colour sums + uint32 count per bucket for bits 32/33 (the reference's
 * `bucket_int`, flam3.c:3772, 3824; adds are truncated and saturate at UINT_MAX,
 * flam3.c:3838-3854), double[4] + double for bits 64 (`bucket_double`, flam3.c:3724). The
 * chaos game does not add into them directly: it adds into float4 staging replicas with one
 * vector reduction per hit, which are exact only below 2^24, so a sweeper warp in every CTA
 * keeps moving whatever has grown past F3B_CARRY_* out of the staging planes into the exact
 * planes while the walkers run, and f3b_hist_merge_kernel moves the rest (f3b_iterate.cu). */
struct F3HistView {
   const void *color; /* uint4[...] or double[4][...] */
   const void *count; /* uint32[...] or double[...]; nullptr: colour channel 3 is the count */
   int f64;           /* bits 64 */
   int pitch;         /* buckets per row */
   int row0, rows;    /* bucket rows [row0, row0 + rows) are held, row0 at element 0 (a strip
                         of the frame reduced over several GPUs; else 0, height) */
};
#ifdef __CUDACC__
__device__ __forceinline__ size_t f3b_view_index(const F3HistView &V, int x, int y) {
   return (size_t)x + (size_t)V.pitch * (size_t)(y - V.row0);
}
__device__ __forceinline__ float4 f3b_view_color(const F3HistView &V, size_t i) {
   if (V.f64) {
      const double4 *p = reinterpret_cast<const double4 *>(V.color);
      const double4 d = p[i];
      return make_float4((float)d.x, (float)d.y, (float)d.z, (float)d.w);
   }
   const uint4 u = reinterpret_cast<const uint4 *>(V.color)[i];
   return make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
}
__device__ __forceinline__ float f3b_view_count(const F3HistView &V, size_t i) {
   if (V.f64) {
      if (V.count) return (float)reinterpret_cast<const double *>(V.count)[i];
      return (float)reinterpret_cast<const double4 *>(V.color)[i].w;
   }
   if (V.count) return (float)reinterpret_cast<const uint32_t *>(V.count)[i];
   return (float)reinterpret_cast<const uint4 *>(V.color)[i].w;
}
#endif
#define F3B_CARRY_U32 4194304.0f /* 2^22: staged integer sums are moved out at this size ... */
#define F3B_EXACT_LIMIT 16777216.0f /* ... and must never reach 2^24 */
#define F3B_CARRY_F64 262144.0f  /* 2^18: bits 64, keeps the float rounding of an add <= 2^-6 */
#define F3B_SWEEP_ENTRIES 256u   /* staging entries a sweeper warp examines per ticket */
#define F3B_SAT_LIST_CAP 65536u
#define F3B_HOT_LIST_CAP 262144u

#define F3B_CHAINS_PER_GROUP 256 /* largest chain group = threads per block of the chaos-game
                                    kernel; batches with few chains per temporal sample use
                                    smaller groups (32..256) so that more of them are resident */

/* Warp-uniform per-sample values served from the kernel parameters (constant bank 0). */
struct F3ConstSample {
   float4 bounds; /* b0 b1 b2 b3       (rect.c:809-812) */
   float4 cs;     /* ws0 hs1 rcx rcy   (rect.c:819-822, rot_center) */
   float4 rot;    /* rot00 rot01 rot10 rot11 */
   uint4 cum;     /* first chooser row when it has <= 4 entries */
   uint4 lay;     /* F3B_HF_* flags | chooser stride << 8 | cmap_size << 20;
                     xform base | stride << 16;  palette offset | chooser offset << 16;
                     offset of the final xform's record (all in blob words) */
   uint4 pick;    /* <= 4 xforms, no xaos: (cumulative count << 18 | record offset) per xform,
                     saturated and padded with the last xform (f3b_iterate.cu, SIMPLE) */
};
#define F3B_HF_XAOS 1u
#define F3B_HF_ROTATE 2u
#define F3B_HF_PLOT_READY 4u
#define F3B_HF_PAL_LINEAR 8u
#define F3B_HF_FINAL 16u
#define F3B_HF_CHOOSER_CONST 32u /* <= 4 xforms, no xaos: thresholds come from samp[].cum */
#define F3B_HF_PAL_PACKED 64u    /* the rgba8 copy of the dmap may be used (F3B_PLOT_PACKED) */
#define F3B_HF_ANY_POST 128u     /* some xform of the sample has a post affine */
/* temporal samples per launch: 296 x 96 B = 27.75 KB of the 32 KB parameter space that CUDA
 * 12.1+ gives sm_70+ kernels. An animation frame with 1000 temporal samples is 4 launches
 * (on side streams, so they overlap). */
#define F3B_PARAM_SAMPLES 296

struct F3IterateParams {
   const uint8_t *blobs;         /* device: packed temporal samples */
   const f3b_sample_job *jobs;   /* device */
   const uint32_t *group_job;    /* device: job index of every chain group */
   uint32_t group_begin;         /* this launch covers chain groups [group_begin, n_groups) */
   uint32_t n_groups;
   uint32_t job_base;            /* job index of samp[0] */
   uint32_t rank, nranks;        /* a rank runs the groups g with g % nranks == rank */
   float4 *hist_color;           /* [height*width] r,g,b,a sums */
   float *hist_count;            /* [height*width] 255*logvis sums; may alias nothing when !count */
   int width, height;
   F3HistLayout lay;             /* bucket (x,y) -> element of the histogram planes */
   int floor_adds;               /* bits 32/33: add floor(value) like the integer buckets */
   uint64_t seed;
   unsigned long long *badvals;  /* device counter */
   /* exact planes and the sweeper warp (threads >= chain_threads of every CTA) */
   void *exact_color;            /* uint4[elems] or double4[elems] */
   void *exact_count;            /* uint32[elems] or double[elems]; nullptr without a count plane */
   uint32_t chain_threads;       /* walkers per CTA = chains per group */
   uint32_t sweep_on;            /* blockDim.x == chain_threads + 32 */
   uint32_t sweep_gap_ns;        /* a cold pass over the staging planes is due this often */
   uint32_t sweep_mode;          /* 1 cold passes | 2 hot list | 4 hot mark from the nominal gap only | 8 walkers report sampled entries (11 = normal; F3B_SWEEP_MODE for experiments) */
   uint32_t exact_f64;
   float carry_thresh;           /* F3B_CARRY_U32 / F3B_CARRY_F64 (F3B_CARRY lowers it, for tests) */
   unsigned int *sweep_state;    /* device: [0] cold ticket, [1] hot entries, [2] hot ticket */
   uint32_t *sweep_hot_list;     /* device: staging entries that grow fast */
   unsigned int *sweep_hot_bitmap; /* device: one bit per staging entry: on the hot list */
   unsigned long long *exact_flags; /* device: [0] staged values seen at >= 2^24 (exactness
                                       lost), [1] entries in sat_list, [4..6] diagnostics: where
                                       the losses were seen (cold pass, hot list, merge) */
   uint32_t *sat_list;           /* device: bucket * 8 + channel of exact sums that wrapped */
   uint32_t warp_uniform;        /* warp-uniform xform step + walker exchange (f3b_iterate.cu) */
   uint32_t wu_smem_off;         /* words: 3 floats per walker of the CTA behind the blob */
   /* binned scatter (histograms that do not fit L2; AGG kernels only, see f3b_iterate.cu):
    * hits become 8-byte records in per-region queues instead of reductions */
   uint32_t bin_regions;         /* 0 = off; else regions of bin_region_rows bucket rows each */
   uint32_t bin_region_rows, bin_row_magic; /* region of row y = umulhi(y, magic) */
   uint32_t bin_cap;             /* blocks of 32 records in the pool */
   uint32_t bin_smem_off, bin_warp_words; /* per-warp queues in shared memory, in words */
   uint2 *bin_records;           /* device: the pool, [bin_cap][32] */
   uint8_t *bin_block_region;    /* device: [bin_cap] region of every block, 0xff = not used */
   unsigned int *bin_pos;        /* device: blocks handed out */
   /* probe launch (AGG kernels only): every chain stores at most probe_iters iterations,
    * nothing is added to the histogram, probe_counts[0] += lanes that reached an add,
    * probe_counts[1] += those that shared their bucket with another lane of the warp */
   uint32_t probe_iters;
   unsigned long long *probe_counts;
   F3ConstSample samp[F3B_PARAM_SAMPLES]; /* parameter-space home of the per-sample constants */
};

/* returns the number of kernel launches made (0 on error; see cudaGetLastError). `consts`
 * holds the constants of jobs [p.job_base, p.job_base + n_consts), n_consts <= capacity */
int f3b_launch_iterate(F3IterateParams &p, const F3ConstSample *consts, int n_consts,
                       bool count_channel, bool simple, bool aggregate, int grid_blocks,
                       int group_size, size_t smem_bytes, cudaStream_t stream);
int f3b_iterate_blocks_per_sm(int group_size, size_t smem_bytes, bool simple);
int f3b_iterate_const_capacity(void);
int f3b_iterate_max_blocks_per_sm(size_t smem_bytes, bool simple);
/* exact planes += all staging copies (saturating for the integer buckets); the staging
 * copies are left zeroed. hot_max receives the fullest bucket's count (uint32, atomicMax). */
struct F3MergeParams {
   float4 *color;
   float *count; /* nullptr without a count plane */
   void *exact_color, *exact_count;
   int f64;
   size_t elems;
   F3HistLayout lay;
   unsigned int *hot_max;
   unsigned long long *exact_flags;
   const uint32_t *sat_list;
   int sat_check; /* sweepers ran: uint32 sums may have wrapped */
};
int f3b_launch_hist_merge(const F3MergeParams &m, cudaStream_t stream);
int f3b_launch_orbits(const uint32_t *blob, size_t blob_bytes, int n_chains, int n_per_chain, int fuse,
                      uint64_t seed, const float *start, float4 *out, unsigned long long *badvals,
                      cudaStream_t stream);
int f3b_launch_debug_apply(const uint32_t *blob, size_t blob_bytes, int xform_index, const float *in,
                           int n, float *out, uint64_t seed, cudaStream_t stream);

struct F3DensityParams {
   F3HistView hist;           /* the exact buckets */
   float4 *accum;
   int width, height;         /* accum / de_kidx are plain [height][width] */
   int row_begin, row_end;    /* target bucket rows written by this launch */
   int oversample;
   float k1, k2;
   int int_accum;             /* bits 32: accumulators are integers (truncate after add) */
   /* density estimation tables (device); n_kernels == 0 -> plain log scale (rect.c:947-974) */
   const float *de_coefs;     /* [n_kernels * kernel_size] octant-packed, filters.c:322-384 */
   const uint8_t *de_radius;  /* [n_kernels] ceil(width)-1, rect.c:123 */
   const float2 *de_kparam;   /* [n_kernels] {4.5 log2(e) / h_k^2, centre coefficient N_k} */
   const uint16_t *de_d2max;  /* [n_kernels] largest dx^2+dy^2 with a non-zero coefficient */
   int de_table_only;         /* F3B_DE_TABLE=1: gather from the coefficient table (A/B, tests) */
   int n_kernels, kernel_size, max_radius;
   int max_filtered_counts, max_filter_index;
   float curve;
   /* scratch planes of the two-pass density estimator (device) */
   uint16_t *de_kidx;         /* [height*width] kernel index per bucket, 0xffff = not a source */
   int *tile_maxrad;          /* [tiles_y*tiles_x] largest radius >= 1 in each 32x32 tile */
   int overwrite_accum;       /* first batch of a frame: write the accumulators, do not add */
   unsigned int *hot_max;     /* device: max over buckets of the count plane */
   /* filled by f3b_launch_density */
   int class_row_begin, class_row_end; /* rows classified: the target rows +- max_radius */
   int class_tile_y0, class_tile_y1, gather_tile_y0;
};
int f3b_launch_density(const F3DensityParams &p, cudaStream_t stream);

struct F3FinalParams {
   float4 *accum; /* modified in place when earlyclip */
   int width, height; /* bucket grid */
   int image_width, image_height, oversample, de_offset;
   const double *filter; /* device [fw*fw] */
   int filter_width;
   int earlyclip, nchan, transp, bytes_per_channel;
   double g, linrange, vibrancy, highpow;
   double background[3];
   void *out;          /* device image, tightly packed image_width*nchan*bpc per row */
   int out_row_begin, out_row_end; /* image rows produced by this launch */
   /* float first, double where it matters (f3b_post.cu): pixels to redo; nullptr = all double */
   uint32_t *fixup_list;
   unsigned int *fixup_count;
};
int f3b_launch_final(const F3FinalParams &p, cudaStream_t stream);

#define F3B_BIN_MAX_REGIONS 16
#define F3B_BIN_NULL 0xffffffffu
/* region queue -> reductions into the L2-resident staging tile of that region */
int f3b_launch_bin_apply(const uint2 *records, const uint8_t *block_region, const unsigned int *pos, uint32_t cap,
                         uint32_t region, float4 *tile_color, float *tile_count, cudaStream_t stream);

/* Several GPUs on one frame: rank r owns a strip of bucket rows and sums every rank's
 * buckets of those rows straight out of the peers' memory (NVLink loads), saturating like
 * the buckets themselves. This is the reduce-scatter by image strip, halo rows included
 * (overlapping strips are simply fetched by both neighbours). */
#define F3B_MAX_RANKS 16
struct F3ReduceParams {
   const void *color[F3B_MAX_RANKS];
   const void *count[F3B_MAX_RANKS]; /* all nullptr without a count plane */
   int npeers, f64;
   void *out_color, *out_count;
   size_t first; /* element of the peers' planes that becomes element 0 of the output */
   size_t n;     /* elements to reduce */
};
int f3b_launch_hist_reduce(const F3ReduceParams &r, cudaStream_t stream);

/* reference bucket layout uint32[5] per bucket <-> the exact planes */
int f3b_launch_hist_from_u32(const uint32_t *src5, void *exact_color, void *exact_count, int f64,
                             int width, int height, int pitch, cudaStream_t stream);
int f3b_launch_hist_to_u32(const void *exact_color, const void *exact_count, int f64, uint32_t *dst5,
                           int width, int height, int pitch, cudaStream_t stream);

#endif
