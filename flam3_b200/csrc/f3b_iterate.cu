/*
 * f3b_iterate.cu — the chaos game + histogram accumulation for sm_100a.
 *
 * Replaces the reference's iter_thread (rect.c:253-527) + flam3_iterate (flam3.c:232-299)
 * + apply_xform (variations.c:2129-2399). Design:
 *
 *  - one persistent walker per thread; a thread runs whole "chains" (= the reference's
 *    sub-batches: random start point, `fuse` discarded iterations, then stored ones);
 *  - a block works on a "chain group": up to 256 chains (32..256, = blockDim.x) of the SAME
 *    temporal sample, whose packed genome/chooser/palette blob is staged once in shared
 *    memory; blocks walk the group list with a grid stride (persistent CTAs). What is
 *    warp-uniform per sample (camera, flags, offsets, the xform pick table) comes from the
 *    kernel parameters through the constant bank instead;
 *  - xform selection reproduces the 16384-grain table exactly through per-row cumulative
 *    counts (flam3.c:165-224), including xaos rows;
 *  - the 5-channel bucket bump (rect.c:500-512) becomes ONE 16-byte vector reduction
 *    `red.global.add.v4.f32` into L2-resident float4 colour buckets (K copies of them, a CTA
 *    adds into copy blockIdx.x % K, see f3b_api.cu: hist_layout_for), plus a scalar
 *    `red.global.add.f32` into the count plane only when the count channel is not
 *    identical to the alpha channel (COUNT=false: palette alpha == 1 and temporal filter
 *    weight == 1 everywhere, the common case, where b[3] == b[4] bucket by bucket);
 *  - inside a staging copy a bucket does not sit at x + pitch * y but at a scrambled place
 *    (f3b_internal.h: F3HistLayout): the hits of an attractor cluster on neighbouring buckets,
 *    and L2 serialises reductions that meet in the same sectors;
 *  - instantiations f3b_iterate_kernel<COUNT, KIND, MODE, WU, SINGLE> instead of run-time
 *    flags - the 72-register loop pays 3-6 % for a few instructions it merely carries:
 *    KIND 1/2 (SIMPLE: classic flames, packed pick, folded flag tests; 2: and no post affine),
 *    MODE 1 (AGG: collapsed attractors, equal buckets of a warp aggregated before the
 *    reduction; also the binned scatter) and 2 (walkers report sampled entries to the sweeper:
 *    histograms beyond L2), WU (warp-uniform xform step for genomes whose xforms run different
 *    variation code), SINGLE (one temporal sample: constants as constant-bank operands).
 *
 * Adds are integer valued (floor, like bump_no_overflow with unsigned buckets,
 * flam3.c:3838-3854), so float accumulation is exact up to 2^24 per channel; sweeper warps move
 * what grows large into the exact uint32 / double buckets while the walkers run (below).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "f3b_internal.h"
#include "f3b_variations.cuh"

#define F3B_BLOCK F3B_CHAINS_PER_GROUP
#ifndef F3B_ITER_MIN_BLOCKS
#define F3B_ITER_MIN_BLOCKS 3 /* 24 walker warps + 3 sweeper warps per SM, 72 registers */
#endif
#ifndef F3B_GENERIC_MIN_BLOCKS
#define F3B_GENERIC_MIN_BLOCKS 2 /* the non-SIMPLE instantiations: 96 registers, 16 walker warps per SM.
                                    With 3 CTAs they have 72 registers, spill, and run 30-47 % slower on
                                    the random genomes of C5 (profiles/r02_generic_occupancy_ab.txt) */
#endif

__device__ __forceinline__ void f3b_red_v4(float4 *addr, float r, float g, float b, float a) {
   asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(r), "f"(g),
                "f"(b), "f"(a)
                : "memory");
}
__device__ __forceinline__ void f3b_red_f32(float *addr, float v) {
   asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

/* ---- the sweeper warp: exact buckets behind float staging planes -------------------- */
/* A float sum of integers is exact only below 2^24 (65 793 hits of 255), and a 1080p
 * quality-2000 frame puts 8e6 hits into its fullest bucket. So the float4 planes the walkers
 * add into are only a STAGING area: the last warp of every CTA runs no chains but keeps
 * walking over the staging planes (all copies, segment after segment from a global ticket
 * counter, so the live sweepers of a launch always cover everything), and whatever has
 * grown to F3B_CARRY_* is moved into the exact planes: an atomic exchange with zero takes a
 * channel out of the staging bucket (what comes back is an exact integer below 2^24 and
 * exactly what left, whatever adds land in between), an integer atomic adds it to the
 * uint32 bucket. A uint32 sum that wraps is recorded and pinned to UINT_MAX
 * after the merge, which is the reference's saturation (flam3.c:3838-3854). A staged value
 * seen at >= 2^24 means the sweep came too late: counted in exact_flags[0], reported by
 * f3b_exactness. Costs no instruction in the walkers' loop. */
__device__ __forceinline__ float4 f3b_ld_staged4(const float4 *p) {
   float4 v;
   asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                : "l"(p));
   return v;
}
__device__ __forceinline__ float f3b_ld_staged1(const float *p) {
   float v;
   asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
   return v;
}
__device__ __noinline__ void f3b_carry_channel(const F3IterateParams &P, size_t bucket, int ch, float v) {
   if (v == 0.0f) return;
   if (P.exact_f64) {
      double *d = ch < 4 ? reinterpret_cast<double *>(P.exact_color) + bucket * 4 + ch
                         : reinterpret_cast<double *>(P.exact_count) + bucket;
      atomicAdd(d, (double)v);
      return;
   }
   unsigned int *u = ch < 4 ? reinterpret_cast<unsigned int *>(P.exact_color) + bucket * 4 + ch
                            : reinterpret_cast<unsigned int *>(P.exact_count) + bucket;
   const unsigned int add = (unsigned int)v; /* integer valued, >= 0 */
   const unsigned int old = atomicAdd(u, add);
   if (old + add < old) { /* wrapped: saturate after the merge */
      const unsigned long long slot = atomicAdd(P.exact_flags + 1, 1ull);
      if (slot < F3B_SAT_LIST_CAP) P.sat_list[slot] = (uint32_t)(bucket * 8 + (size_t)ch);
   }
}
/* Warp-level aggregation of bucket adds, for attractors that collapse onto a few buckets:
 * reductions to ONE address serialise in L2 (~2e9/s for the whole GPU), so a genome whose
 * orbit sits on a fixed point or a short cycle ran 50x slower than a spread-out one
 * (profiles/r01_c5_*: var76 267 -> 39 ms, var47 118 -> 22 ms, var58 175 -> 35 ms). In the AGG
 * instantiations of the kernel the lanes that are about to add compare their buckets
 * (match.any); lanes with the same bucket sum their colours through shuffles and the lowest
 * one adds the sum, with integer atomics, directly to the exact bucket. The sums are of
 * integer-valued floats, so they stay exact. Returns false for a lane whose add has been
 * taken care of (folded into another lane's, or made by this function). The check costs ~5 % on
 * a spread-out flame, so the host only launches the AGG kernels when a short probe launch of
 * the same kernel (probe_counts != NULL: count, add nothing) found the lanes colliding.
 * Every warp-synchronous operation below is executed by exactly the lanes named in its mask:
 * match and ballot by all lanes that entered together (`act`), the shuffles by the lanes of
 * one `peers` set, which all see the same set and so run the same number of rounds. */
__device__ __noinline__ bool f3b_bump_aggregate(const F3IterateParams &P, size_t bucket,
                                                unsigned long long bidx, float4 *colp, float *cntp,
                                                unsigned long long *probe_counts) {
   const unsigned act = __activemask();
   const unsigned lane = threadIdx.x & 31u;
   const unsigned peers = __match_any_sync(act, bidx);
   const bool dup = (peers & (peers - 1u)) != 0u;
   const unsigned nd = __popc(__ballot_sync(act, dup));
   if (probe_counts) { /* probe launch: measure, add nothing */
      if ((int)lane == __ffs(act) - 1) {
         atomicAdd(probe_counts, (unsigned long long)__popc(act));
         atomicAdd(probe_counts + 1, (unsigned long long)nd);
      }
      return false;
   }
   if (!dup) return true;
   const int leader = __ffs(peers) - 1;
   const float4 a = *colp;
   const float c = *cntp;
   float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f, sc = 0.f;
   for (unsigned rest = peers & ~(1u << leader); rest; rest &= rest - 1u) {
      const int src = __ffs(rest) - 1;
      sx += __shfl_sync(peers, a.x, src);
      sy += __shfl_sync(peers, a.y, src);
      sz += __shfl_sync(peers, a.z, src);
      sw += __shfl_sync(peers, a.w, src);
      sc += __shfl_sync(peers, c, src);
   }
   if ((int)lane != leader) return false;
   /* A bucket that several lanes of one warp hit at once is a bucket that fills faster than
    * any sweeper could empty its staging copy (a collapsed attractor adds 1e12 per second to
    * one address): the aggregated sum goes straight into the exact bucket. Lone hits of the
    * same bucket still take the staged path, at a rate the sweepers handle. */
   const float tx = a.x + sx, ty = a.y + sy, tz = a.z + sz, tw = a.w + sw, tc = c + sc;
   if (!P.exact_f64) {
      /* The count channel (the alpha channel when it doubles as the count) is never smaller
       * than the others: every hit adds 255 vis to it and at most that to the rest. So ONE
       * atomic with a return value - on the count - tells whether anything can wrap: while it
       * is below 2^31 the other channels go out as plain reductions (r and g as one 64-bit
       * reduction: no carry can cross the halves), which is what keeps the hot addresses of a
       * collapsed attractor from serialising on four round trips per add. Near saturation
       * every channel takes the careful path. */
      unsigned int *ec = reinterpret_cast<unsigned int *>(P.exact_color) + bucket * 4;
      unsigned int *top = P.hist_count ? reinterpret_cast<unsigned int *>(P.exact_count) + bucket : ec + 3;
      const unsigned int addt = (unsigned int)(P.hist_count ? tc : tw);
      const unsigned int old = atomicAdd(top, addt);
      if (old < 0x7fff0000u && old + addt < 0x80000000u) {
         const unsigned long long rg = (unsigned long long)(unsigned int)tx | ((unsigned long long)(unsigned int)ty << 32);
         asm volatile("red.global.add.u64 [%0], %1;" ::"l"(ec), "l"(rg) : "memory");
         asm volatile("red.global.add.u32 [%0], %1;" ::"l"(ec + 2), "r"((unsigned int)tz) : "memory");
         if (P.hist_count) asm volatile("red.global.add.u32 [%0], %1;" ::"l"(ec + 3), "r"((unsigned int)tw) : "memory");
         return false;
      }
      if (old + addt < old) { /* the count itself wrapped: pinned to UINT_MAX after the merge */
         const unsigned long long slot = atomicAdd(P.exact_flags + 1, 1ull);
         if (slot < F3B_SAT_LIST_CAP) P.sat_list[slot] = (uint32_t)(bucket * 8 + (size_t)(P.hist_count ? 4 : 3));
      }
      f3b_carry_channel(P, bucket, 0, tx);
      f3b_carry_channel(P, bucket, 1, ty);
      f3b_carry_channel(P, bucket, 2, tz);
      if (P.hist_count) f3b_carry_channel(P, bucket, 3, tw);
      return false;
   }
   f3b_carry_channel(P, bucket, 0, tx);
   f3b_carry_channel(P, bucket, 1, ty);
   f3b_carry_channel(P, bucket, 2, tz);
   f3b_carry_channel(P, bucket, 3, tw);
   if (P.hist_count) f3b_carry_channel(P, bucket, 4, tc);
   return false;
}

/* ---- binned scatter: histograms larger than L2 --------------------------------------- */
/* A 4K frame at oversample 2 has 33.6 million buckets (538 MB of float4): scattered
 * reductions into it are DRAM read-modify-writes of a 32-byte sector per hit and run 3.7x
 * slower than into an L2-resident plane. So the hit is not applied where it is produced: it
 * becomes an 8-byte record {bucket within its region | count << 24, r g b a bytes} (the adds
 * of the integer buckets are whole numbers <= 255) destined for its REGION, a band of bucket
 * rows whose float4 staging tile fits L2. Every warp keeps a 64-record queue per region in
 * shared memory; whenever one holds 32 records they go out as one 256-byte block of a global
 * pool (blocks are handed to warps 32 at a time from one counter; a byte per block says
 * which region it belongs to): 8 B per hit of sequential DRAM writes instead of 64 B of
 * scattered sector traffic, and the pool only needs one record per iteration of the launch,
 * however the hits spread over the regions. After the launch the regions are applied one
 * after the other (f3b_bin_apply_kernel: the blocks of region r, one vector reduction per
 * record into the tile, which stays in L2) and each tile is merged into the exact buckets of
 * its region.
 * f3b_bin_commit is called by the whole warp at a converged point of the iteration loop (the
 * queues are warp-private state: lanes arriving in separate groups would race on them). */
#define F3B_BIN_HDR 32u /* words before the queues: [0..16) records queued per region */
/* next block of the pool for this warp (warp-uniform state in registers: blocks are handed
 * out 32 at a time from the global counter) */
__device__ __forceinline__ uint32_t f3b_bin_block(unsigned int *bin_pos, uint32_t &blk_next, uint32_t &blk_end,
                                                  unsigned lane) {
   if (blk_next == blk_end) {
      uint32_t n = 0;
      if (lane == 0) n = atomicAdd(bin_pos, 32u);
      blk_next = __shfl_sync(0xffffffffu, n, 0);
      blk_end = blk_next + 32u;
   }
   return blk_next++;
}
/* inlined into the kernel: P's fields then come straight from the constant bank */
__device__ __forceinline__ void f3b_bin_commit(const F3IterateParams &P, uint32_t *wq, uint32_t a, uint32_t c,
                                               uint32_t r, uint32_t &blk_next, uint32_t &blk_end) {
   const unsigned lane = threadIdx.x & 31u;
   const bool has = a != F3B_BIN_NULL;
   const unsigned valid = __ballot_sync(0xffffffffu, has);
   if (!valid) return;
   uint2 *q = reinterpret_cast<uint2 *>(wq + F3B_BIN_HDR);
   uint32_t newc = 0;
   bool full = false;
   if (has) {
      const unsigned peers = __match_any_sync(valid, r);
      const uint32_t base = *(volatile uint32_t *)(wq + r);
      q[r * 65u + base + __popc(peers & ((1u << lane) - 1u))] = make_uint2(a, c);
      if ((int)lane == __ffs(peers) - 1) {
         newc = base + __popc(peers);
         full = newc >= 32u;
      }
   }
   __syncwarp();
   if (newc && !full) wq[r] = newc;
   for (unsigned fm = __ballot_sync(0xffffffffu, full); fm; fm &= fm - 1u) {
      const int src = __ffs(fm) - 1;
      const uint32_t rr = __shfl_sync(0xffffffffu, r, src);
      const uint32_t nc = __shfl_sync(0xffffffffu, newc, src);
      const uint32_t blk = f3b_bin_block(P.bin_pos, blk_next, blk_end, lane);
      uint2 *qq = q + rr * 65u;
      const uint2 out = qq[lane];
      const uint2 rest = qq[32u + lane];
      if (blk < P.bin_cap) {
         P.bin_records[(size_t)blk * 32u + lane] = out;
         if (lane == 0) P.bin_block_region[blk] = (uint8_t)rr;
      } else if (lane == 0)
         atomicAdd(P.exact_flags, 32ull); /* cannot happen with the host's sizing; never silent */
      __syncwarp();
      if (lane < nc - 32u) qq[lane] = rest;
      if (lane == 0) wq[rr] = nc - 32u;
      __syncwarp();
   }
}
/* end of a CTA's work: what is left in the warp's queues goes out, padded to 32 records */
__device__ __noinline__ void f3b_bin_flush(const F3IterateParams &P, uint32_t *wq, uint32_t blk_next,
                                           uint32_t blk_end) {
   const unsigned lane = threadIdx.x & 31u;
   const uint2 *q = reinterpret_cast<const uint2 *>(wq + F3B_BIN_HDR);
   __syncwarp();
   for (uint32_t r = 0; r < P.bin_regions; r++) {
      const uint32_t c = *(volatile uint32_t *)(wq + r);
      if (c == 0) continue;
      const uint32_t blk = f3b_bin_block(P.bin_pos, blk_next, blk_end, lane);
      const uint2 rec = lane < c ? q[r * 65u + lane] : make_uint2(F3B_BIN_NULL, 0u);
      if (blk < P.bin_cap) {
         P.bin_records[(size_t)blk * 32u + lane] = rec;
         if (lane == 0) P.bin_block_region[blk] = (uint8_t)r;
      } else if (lane == 0)
         atomicAdd(P.exact_flags, 32ull);
      __syncwarp();
      if (lane == 0) wq[r] = 0;
   }
   __syncwarp();
}

/* the blocks of one region -> reductions into the region's L2-resident tile: a warp looks at
 * 32 table entries at a time and handles the blocks that are the region's */
__global__ void __launch_bounds__(256) f3b_bin_apply_kernel(const uint2 *records, const uint8_t *block_region,
                                                            const unsigned int *pos, uint32_t cap, uint32_t region,
                                                            float4 *tile_color, float *tile_count) {
   uint32_t nblk = *pos;
   if (nblk > cap) nblk = cap;
   const unsigned lane = threadIdx.x & 31u;
   const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
   for (uint32_t b0 = warp * 32u; b0 < nblk; b0 += nwarps * 32u) {
      const bool is_mine = b0 + lane < nblk && block_region[b0 + lane] == region;
      for (unsigned m = __ballot_sync(0xffffffffu, is_mine); m; m &= m - 1u) {
         const uint2 rec = records[(size_t)(b0 + (uint32_t)__ffs(m) - 1u) * 32u + lane];
         if (rec.x == F3B_BIN_NULL) continue;
         const uint32_t idx = rec.x & 0xffffffu, c = rec.y;
         f3b_red_v4(tile_color + idx, (float)(c & 255u), (float)((c >> 8) & 255u), (float)((c >> 16) & 255u),
                    (float)(c >> 24));
         if (tile_count) f3b_red_f32(tile_count + idx, (float)(rec.x >> 24));
      }
   }
}
int f3b_launch_bin_apply(const uint2 *records, const uint8_t *block_region, const unsigned int *pos, uint32_t cap,
                         uint32_t region, float4 *tile_color, float *tile_count, cudaStream_t stream) {
   f3b_bin_apply_kernel<<<148 * 8, 256, 0, stream>>>(records, block_region, pos, cap, region, tile_color, tile_count);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}

/* examine one staging entry: carry what has reached the threshold; returns the largest
 * channel seen (before the carry) */
__device__ __forceinline__ float f3b_sweep_entry(const F3IterateParams &P, size_t e, float4 v, float c,
                                                 int where) {
   const float thr = P.carry_thresh;
   const float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
   const size_t bucket = f3b_stage_bucket(P.lay, (uint32_t)(e % P.lay.replica_stride));
   if (m >= thr) {
      /* take the channels out one by one with an exchange: what comes back is exactly what
       * left the staging bucket, whatever lands there meanwhile */
      float *f = reinterpret_cast<float *>(P.hist_color + e);
#pragma unroll
      for (int ch = 0; ch < 4; ch++) {
         const float take = atomicExch(f + ch, 0.0f);
         if (!P.exact_f64 && fabsf(take) >= F3B_EXACT_LIMIT) {
            atomicAdd(P.exact_flags, 1ull);
            atomicAdd(P.exact_flags + 4 + where, 1ull); /* diagnostics: seen by a cold pass / hot service */
         }
         f3b_carry_channel(P, bucket, ch, take);
      }
   }
   if (fabsf(c) >= thr) {
      const float take = atomicExch(P.hist_count + e, 0.0f);
      if (!P.exact_f64 && fabsf(take) >= F3B_EXACT_LIMIT) {
         atomicAdd(P.exact_flags, 1ull);
         atomicAdd(P.exact_flags + 4 + where, 1ull);
      }
      f3b_carry_channel(P, bucket, 4, take);
   }
   return fmaxf(m, fabsf(c));
}
__device__ __forceinline__ unsigned long long f3b_now_ns() {
   unsigned long long t;
   asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
   return t;
}
/* Two kinds of work, so that exactness costs (almost) no L2 bandwidth:
 *  - COLD passes over all staging entries, 128 per ticket from a global counter (whichever
 *    sweepers are alive share a pass). Passes 0 and 1 run flat out from the start of the
 *    launch (~70 us each for a 1080p frame: the fullest bucket of a flame that does not
 *    qualify for warp aggregation needs >= 200 us to reach 2^24), the later ones are due 4x
 *    later each up to one per P.sweep_gap_ns: an entry that was still small when last seen
 *    needs proportionally long to become large.
 *  - Entries that fill fast enough to reach the carry threshold within one such gap go on the
 *    HOT list (once: a bitmap), and between cold tickets every sweeper keeps servicing 32
 *    hot entries at a time. A 1080p
 *    quality-2000 flame has ~6000 hot entries among 6.4 million; they are what actually needs
 *    carrying, every ~100 us for the fullest. */
/* ring of staging entries the walkers report (see the walkers' loop): written by them, read by
 * the CTA's sweeper warp; a lost or repeated report is harmless */
#define F3B_SAMPLE_RING 256u
#define F3B_SAMPLE_EVERY 512 /* a walker reports the entry it just added to every so many iterations */
struct F3SampleRing {
   unsigned int head;
   uint32_t entry[F3B_SAMPLE_RING];
};
__device__ __noinline__ void f3b_sweeper(const F3IterateParams &P, volatile int *s_done, volatile F3SampleRing *ring) {
   const unsigned lane = threadIdx.x & 31u;
   const size_t n_entries = P.lay.replica_stride * P.lay.replicas;
   const unsigned nseg = (unsigned)((n_entries + F3B_SWEEP_ENTRIES - 1) / F3B_SWEEP_ENTRIES);
   const float gap_inv = 1.0f / (float)P.sweep_gap_ns;
   constexpr int U = F3B_SWEEP_ENTRIES / 32;
   unsigned int *const S = P.sweep_state; /* [0] cold ticket, [1] hot entries, [2] hot ticket */
   const unsigned long long t0 = f3b_now_ns();
   if (!(P.sweep_mode & 3u)) return; /* experiments: idle sweepers */
   /* the ticket after the one in hand is requested before the one in hand is worked on */
   unsigned next = 0;
   if (lane == 0) next = atomicAdd(S, 1u);
   unsigned ticket = 0;
   bool have_ticket = false;
   unsigned ring_tail = 0;
   unsigned my_pass = 0; /* the last pass this warp held a ticket of, when, and the measured period */
   unsigned long long my_pass_t = 0;
   float period_ns = 0.0f;
   for (unsigned round = 0;; round++) {
      int done = 0; /* (atomics on both sides of the flag and of the ring: nothing for racecheck to report) */
      if (lane == 0) done = atomicAdd((int *)s_done, 0);
      if (__shfl_sync(0xffffffffu, done, 0)) break;
      bool worked = false;
      if (!have_ticket) {
         ticket = __shfl_sync(0xffffffffu, next, 0);
         if (lane == 0) next = atomicAdd(S, 1u);
         have_ticket = true;
      }
      const unsigned pass = ticket / nseg;
      /* due times: 0, 0, 100 us, 400 us, then one pass per sweep_gap_ns */
      unsigned long long due = 0;
      if (pass == 2) due = 100000ull;
      else if (pass == 3) due = 400000ull;
      else if (pass > 3) due = 400000ull + (unsigned long long)(pass - 3) * P.sweep_gap_ns;
      const unsigned long long elapsed = f3b_now_ns() - t0;
      if ((P.sweep_mode & 1u) && (pass < 2 || elapsed >= due)) {
         /* HOT means: fills at a rate that reaches the carry threshold within one gap between
          * cold passes. Whatever is below that rate cannot get from the threshold of this visit
          * past twice the carry threshold before the next one. */
         /* The gap is what the schedule asks for; what counts is how long the passes really
          * take: with a histogram that lives in DRAM the walkers' scatter saturates the memory
          * system and a pass over the staging planes can take several gaps. The time between
          * this warp's first tickets of consecutive passes is that period. An entry left alone
          * at this visit fills at less than thresh / T, so it stays below 2^24 = 4 thresh as long
          * as the next visit comes within 3 T. */
         if (pass != my_pass) {
            period_ns = (float)(elapsed - my_pass_t) / (float)(pass - my_pass);
            my_pass = pass;
            my_pass_t = elapsed;
         }
         float frac = (float)elapsed * gap_inv;
         if (!(P.sweep_mode & 4u) && period_ns * gap_inv > 1.0f) frac = (float)elapsed / period_ns;
         const float hot_mark = P.carry_thresh * (frac < 1.0f ? fmaxf(frac, 1.0f / 4096.0f) : 1.0f);
         const size_t base = (size_t)(ticket % nseg) * F3B_SWEEP_ENTRIES;
         float4 v[U];
         float c[U];
#pragma unroll
         for (int u = 0; u < U; u++) {
            const size_t e = base + (size_t)u * 32 + lane;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            c[u] = 0.f;
            if (e < n_entries) {
               v[u] = f3b_ld_staged4(P.hist_color + e);
               if (P.hist_count) c[u] = f3b_ld_staged1(P.hist_count + e);
            }
         }
#pragma unroll
         for (int u = 0; u < U; u++) {
            const size_t e = base + (size_t)u * 32 + lane;
            const float m = fmaxf(fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))),
                                  fabsf(c[u]));
            if (m >= hot_mark) {
               f3b_sweep_entry(P, e, v[u], c[u], 0);
               const unsigned bit = 1u << (e & 31u);
               const unsigned old = atomicOr(P.sweep_hot_bitmap + (e >> 5), bit);
               if (!(old & bit)) {
                  const unsigned slot = atomicAdd(S + 1, 1u);
                  if (slot < F3B_HOT_LIST_CAP) P.sweep_hot_list[slot] = (uint32_t)e;
               }
            }
         }
         have_ticket = false;
         worked = true;
         if (pass < 2 && (round & 3u)) continue; /* first passes: as fast as possible */
      }
      /* Entries the walkers have just added to, one in F3B_SAMPLE_EVERY adds: a sample of the
       * buckets weighted by how fast they fill. Whatever is found at the carry threshold is
       * carried now and goes on the hot list. This is what finds the fast fillers of a
       * histogram that lives in DRAM, where a cold pass over the staging planes takes
       * several ms beside the walkers' scatter (4K oversample 2: 7.7 ms, in which the fullest
       * buckets of a still pass 2^24): such a bucket is reported ~100 times on its way from
       * the threshold to 2^24, however long the cold passes take. */
      if (P.sweep_mode & 8u) {
         const unsigned head = atomicAdd((unsigned int *)&ring->head, 0u);
         if (head - ring_tail > F3B_SAMPLE_RING) ring_tail = head - F3B_SAMPLE_RING;
         for (; ring_tail != head; ring_tail += (head - ring_tail < 32u ? head - ring_tail : 32u)) {
            if (lane < head - ring_tail) {
               const size_t e = atomicAdd((uint32_t *)&ring->entry[(ring_tail + lane) % F3B_SAMPLE_RING], 0u);
               if (e < n_entries) {
                  const float4 v = f3b_ld_staged4(P.hist_color + e);
                  const float c = P.hist_count ? f3b_ld_staged1(P.hist_count + e) : 0.0f;
                  if (f3b_sweep_entry(P, e, v, c, 1) >= P.carry_thresh) {
                     const unsigned bit = 1u << (e & 31u);
                     const unsigned old = atomicOr(P.sweep_hot_bitmap + (e >> 5), bit);
                     if (!(old & bit)) {
                        const unsigned slot = atomicAdd(S + 1, 1u);
                        if (slot < F3B_HOT_LIST_CAP) P.sweep_hot_list[slot] = (uint32_t)e;
                     }
                  }
               }
            }
         }
      }
      /* hot entries */
      unsigned nhot = 0;
      if (lane == 0) nhot = *(volatile unsigned int *)(S + 1);
      nhot = __shfl_sync(0xffffffffu, nhot, 0);
      if (nhot > F3B_HOT_LIST_CAP) nhot = F3B_HOT_LIST_CAP;
      if (!(P.sweep_mode & 2u)) nhot = 0;
      if (nhot) {
         unsigned h = 0;
         if (lane == 0) h = atomicAdd(S + 2, 1u);
         h = __shfl_sync(0xffffffffu, h, 0);
         const unsigned groups = (nhot + 31u) / 32u;
         const unsigned idx = (h % groups) * 32u + lane;
         if (idx < nhot) {
            const size_t e = *(volatile uint32_t *)(P.sweep_hot_list + idx);
            if (e < n_entries) { /* a slot may be claimed but not written yet */
               const float4 v = f3b_ld_staged4(P.hist_color + e);
               const float c = P.hist_count ? f3b_ld_staged1(P.hist_count + e) : 0.0f;
               f3b_sweep_entry(P, e, v, c, 1);
            }
         }
      }
      /* every hot entry is looked at about every 20 us by the sweepers of the launch together */
      if (!worked) {
         const unsigned groups = (nhot + 31u) / 32u;
         unsigned ns = groups ? 20000u * gridDim.x / groups : 20000u;
         __nanosleep(ns < 1000u ? 1000u : (ns > 50000u ? 50000u : ns));
      }
   }
}
/* barrier among the walkers only (the sweeper warp never joins it) */
__device__ __forceinline__ void f3b_chain_barrier(uint32_t nthreads) {
   asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

/* The final xform (flam3.c:280-287) is applied out of line: few genomes have one, and a
 * second inlined copy of the variation interpreter would double the hot loop's code. */
__device__ __noinline__ float4 f3b_apply_final(const float *sf, uint32_t xo, float px, float py,
                                               float pc, F3Rng *rngp) {
   F3Rng rng = *rngp;
   float fx, fy, fc, fv;
   f3b_apply_xform(sf, xo, px, py, pc, fx, fy, fc, fv, rng);
   *rngp = rng;
   return make_float4(fx, fy, fc, fv);
}

/* per-sample constants folded into one register */
#define HF_XAOS F3B_HF_XAOS
#define HF_ROTATE F3B_HF_ROTATE
#define HF_PLOT_READY F3B_HF_PLOT_READY
#define HF_PAL_LINEAR F3B_HF_PAL_LINEAR
#define HF_FINAL F3B_HF_FINAL
#define HF_CHOOSER_CONST F3B_HF_CHOOSER_CONST
#define HF_PAL_PACKED F3B_HF_PAL_PACKED

/* Warp-uniform per-sample values (camera, small chooser row) travel in the kernel
 * parameters and are served by the constant cache: every lane needs them every iteration,
 * and fetching them through shared memory costs L1TEX data-pipe bandwidth per lane
 * (128 B/clk/SM however uniform the address), which is what bounds this kernel
 * (profiles/r01_v2: 94% of peak). A __constant__ array was tried and dropped: its speed
 * depended on the module loading mode (31 ms lazy vs 23 ms eager for the same code). */
#define c_samp(ci) P.samp[ci]



/* SIMPLE: every temporal sample of the launch has <= 4 xforms, no xaos, no final xform, a
 * step palette and fully visible xforms (the classic flame): the flag tests fold away and
 * the xform pick is three compares on a packed table from the constant bank. */
template <bool COUNT, int KIND, int MODE, bool WU, bool SINGLE>
__global__ void __launch_bounds__(F3B_BLOCK + 32, KIND ? F3B_ITER_MIN_BLOCKS : F3B_GENERIC_MIN_BLOCKS)
f3b_iterate_kernel(const __grid_constant__ F3IterateParams P) {
   /* MODE 0: one vector reduction per hit; 1 (AGG): lanes of a warp that hit the same bucket add
    * once (also the binned scatter); 2 (REPORT): as 0, and the walkers report a sample of the
    * entries they add to (histograms in DRAM). Instantiations, because merely carrying the few
    * instructions of 2 costs the 72-register loop of 0 six per cent. */
   /* KIND 0: any genome; 1 (SIMPLE): the classic flame, see above; 2: SIMPLE and no xform has a
    * post affine - if-converted, the post block costs every iteration 8 issue slots (-4 %) */
   constexpr bool SIMPLE = KIND != 0;
   constexpr bool NOPOST = KIND == 2;
   constexpr bool AGG = MODE == 1;
   constexpr bool REPORT = MODE == 2;
   /* SINGLE: the launch has one temporal sample (every still). Its camera, flags and pick table
    * are then at fixed offsets of the parameter block and become constant-bank operands of the
    * instructions that use them, instead of ~20 indexed constant loads and their address
    * arithmetic per iteration (of ~143 instructions). */
   extern __shared__ __align__(16) uint32_t smem[];
   const float *sf = reinterpret_cast<const float *>(smem);
   __shared__ int s_done;
   __shared__ F3SampleRing s_ring;

   const uint32_t tid = threadIdx.x;
   const uint32_t nct = P.chain_threads; /* walkers; one more warp sweeps when P.sweep_on */
   if (P.sweep_on) {
      if (tid == 0) {
         s_done = 0;
         s_ring.head = 0;
      }
      __syncthreads();
      if (tid >= nct) {
         f3b_sweeper(P, &s_done, &s_ring);
         return;
      }
   }
   const uint32_t rep_off = (blockIdx.x % P.lay.replicas) * (uint32_t)P.lay.replica_stride;
   uint32_t *const wq = smem + P.bin_smem_off + (tid >> 5) * P.bin_warp_words; /* binned scatter only */
   uint32_t blk_next = 0, blk_end = 0; /* no pool blocks in hand */
   if (AGG && P.bin_regions) { /* this warp's region queues start empty */
      wq[tid & 31u] = 0u;
      __syncwarp();
   }
   uint32_t loaded_job = 0xffffffffu;
   unsigned int bad = 0;

   /* this launch covers groups [group_begin, n_groups); a rank runs the groups g with
    * g % nranks == rank: local index lg -> g = first_g + lg * nranks */
   const uint32_t first_g =
       P.group_begin + (P.rank + P.nranks - P.group_begin % P.nranks) % P.nranks;
   for (uint32_t lg = blockIdx.x;; lg += gridDim.x) {
      const uint64_t g64 = (uint64_t)first_g + (uint64_t)lg * P.nranks;
      if (g64 >= P.n_groups) break;
      const uint32_t g = (uint32_t)g64;
      const uint32_t ji = P.group_job[g];
      const f3b_sample_job job = P.jobs[ji];

      if (ji != loaded_job) {
         f3b_chain_barrier(nct); /* everyone is done with the previous blob */
         const uint32_t *src = reinterpret_cast<const uint32_t *>(P.blobs + job.blob_offset);
         const uint32_t words = src[F3B_H_WORDS];
         for (uint32_t i = tid; i < words; i += nct) smem[i] = src[i];
         f3b_chain_barrier(nct);
         loaded_job = ji;
      }

      const uint32_t chain = (g - job.first_group) * nct + tid; /* nct chains per group */
      const bool binned = AGG && P.bin_regions != 0;
      /* Warp-uniform xform step: the 32 walkers of a warp apply the SAME xform in an iteration
       * (picked from one lane's draw, a different lane every iteration). Each walker still
       * sees xforms that are independent over time with the chooser's probabilities, so its
       * orbit is a chaos game as before - but the three LDS.128 of the xform record become
       * broadcasts (12 -> 3 wavefronts of the ~41 per warp-iteration that bound the kernel)
       * and the lanes run the same variation code (random genomes: 6-16 of 32 lanes active
       * otherwise). Walkers that kept sharing their xforms would contract onto one another,
       * so every 4 iterations the CTA's walkers are dealt out anew: thread (warp w, lane l)
       * takes over the walker of thread (l * W + w) of W warps - each warp's 32 walkers come
       * from all the others. Not with xaos (the pick depends on the walker's last xform), not
       * for CTAs of fewer than 4 warps. A template parameter: merely carrying the code cost the
       * kernels that do not use it 6 %. */
      const bool wu = WU && !binned && !(c_samp(SINGLE ? 0u : ji - P.job_base).lay.x & F3B_HF_XAOS);
      const bool lockstep = binned || wu;
      if (chain >= job.n_chains && !lockstep) continue;

      /* stored iterations of this chain */
      const uint64_t base = job.n_iters / job.n_chains;
      const uint32_t rem = (uint32_t)(job.n_iters % job.n_chains);
      int n_store = chain < job.n_chains ? (int)(base + (chain < rem ? 1u : 0u)) : 0;
      if (n_store == 0 && !lockstep) continue;
      if (AGG && P.probe_iters && n_store > (int)P.probe_iters) n_store = (int)P.probe_iters;
      /* binned scatter: the lanes of a warp stay together through the loop (the records are
       * committed by the whole warp), a lane with fewer iterations idles at the end; the
       * warp-uniform step keeps the whole CTA together (the walkers change hands) */
      int n_loop = n_store;
      if (binned) {
         n_loop = __reduce_max_sync(0xffffffffu, n_store);
         if (n_loop == 0) continue;
      } else if (wu) {
         n_loop = (int)(base + (rem ? 1u : 0u));
         if (AGG && P.probe_iters && n_loop > (int)P.probe_iters) n_loop = (int)P.probe_iters;
         if (n_loop == 0) continue;
      }

      /* Flags, chooser stride and palette size share one register (made opaque, or the
       * compiler splits it into a register per flag and spills them); the blob offsets are
       * fetched from the constant bank where they are used. */
      uint32_t ci = SINGLE ? 0u : ji - P.job_base; /* index into c_samp[] */
      uint32_t hf = c_samp(ci).lay.x;
      asm volatile("" : "+r"(hf));

      F3Rng rng;
      f3b_rng_seed(rng, P.seed, job.sample_id, chain);

      /* rect.c:394-397: start point in [-1,1]^2, colour in [0,1]; the 4th draw seeds an
       * opacity that is overwritten before first use, it is drawn to keep the pattern */
      float px = rng.u11();
      float py = rng.u11();
      float pc = rng.u01();
      (void)rng.u01();
      uint32_t lastxf = 0;

      for (int i = -(int)job.fuse; i < n_loop; i++) {
         uint32_t bin_a = F3B_BIN_NULL, bin_c = 0, bin_r = 0;
         uint32_t wu_rnd = 0;
         if (wu) { /* every lane draws, lane (i mod 32)'s draw picks the xform for all */
            wu_rnd = rng.next();
            wu_rnd = __shfl_sync(0xffffffffu, wu_rnd, (unsigned)(i + 1024) & 31u);
         }
         do { /* `continue` below ends the iteration's work and falls through to the commit */
         /* a lane without (more) iterations of its own idles; with the warp-uniform step it
          * still runs the fuse iterations, so that the walker it holds is on the attractor
          * when another thread takes it over */
         if (lockstep && i >= n_store && (binned || i >= 0)) continue;
         float qx, qy, qc, qv;
         uint32_t fn = 0;
         /* flam3.c:249-270: pick + apply, retrying a bad value up to 4 times in place */
         for (int consec = 0;;) {
            const uint32_t rnd = (wu && consec == 0) ? wu_rnd : rng.next();
            const uint32_t k = rnd & (F3B_CHOOSER_GRAIN - 1u);
            /* `ci` is redefined opaquely so the constant loads below stay in the loop; hoisted,
             * they would occupy a dozen registers across the variation code */
            if (!SINGLE) asm volatile("" : "+r"(ci));
            uint32_t xo;
            if (SIMPLE) {
               /* (k << 18 | 0x3ffff) against (count << 18 | record offset) */
               const uint32_t key = rnd * 262144u + 0x3ffffu; /* one multiply-add */
               const uint4 e = c_samp(ci).pick;
               /* the entries ascend: selects from the last to the first leave the first one
                * that exceeds the key - no branches for the lanes to diverge on */
               uint32_t sel = e.w;
               sel = key < e.z ? e.z : sel;
               sel = key < e.y ? e.y : sel;
               sel = key < e.x ? e.x : sel;
               xo = sel & 0x3ffffu;
            } else {
            /* fn = number of cumulative thresholds <= k; rows are padded with 16384 */
            if (hf & HF_CHOOSER_CONST) {
               const uint4 c = c_samp(ci).cum;
               fn = (k >= c.x ? 1u : 0u) + (k >= c.y ? 1u : 0u) + (k >= c.z ? 1u : 0u);
            } else {
               const uint32_t cstride = (hf >> 8) & 0xfffu;
               const uint32_t ro = (c_samp(ci).lay.z >> 16) + ((hf & HF_XAOS) ? lastxf * cstride : 0u);
               fn = 0;
               for (uint32_t j = 0; j < cstride; j += 4) {
                  const uint4 c = *reinterpret_cast<const uint4 *>(smem + ro + j);
                  fn += (k >= c.x ? 1u : 0u) + (k >= c.y ? 1u : 0u) + (k >= c.z ? 1u : 0u) +
                        (k >= c.w ? 1u : 0u);
               }
            }
            const uint32_t xfl = c_samp(ci).lay.y; /* base | stride << 16 */
            xo = (xfl & 0xffffu) + fn * (xfl >> 16);
            }
            const bool isbad = f3b_apply_xform(sf, xo, px, py, pc, qx, qy, qc, qv, rng,
                                               !NOPOST);
            if (!isbad) break;
            bad++;
            if (++consec >= 5) break;
            px = qx;
            py = qy;
            pc = qc;
         }
         lastxf = fn + 1;
         px = qx;
         py = qy;
         pc = qc;

         /* flam3.c:280-287: final xform does not feed back into the orbit */
         if (!SIMPLE && (hf & HF_FINAL)) {
            const float final_opacity = sf[F3B_H_FINAL_OPACITY];
            bool apply = (final_opacity == 1.0f);
            if (!apply) apply = rng.u01() < final_opacity;
            if (apply) {
               F3Rng tmp = rng;
               const float4 f =
                   f3b_apply_final(sf, c_samp(ci).lay.w, px, py, pc, &tmp);
               rng = tmp;
               qx = f.x;
               qy = f.y;
               qc = f.z; /* opacity stays that of the regular xform */
            }
         }

         if (i < 0) continue;

         /* rect.c:419-441: camera */
         if (!SINGLE) asm volatile("" : "+r"(ci));
         const float4 bb = c_samp(ci).bounds; /* b0 b1 b2 b3 */
         const float4 cs = c_samp(ci).cs;     /* ws0 hs1 rcx rcy */
         float x = qx, y = qy;
         if (hf & HF_ROTATE) {
            const float4 rot = c_samp(ci).rot;
            const float dx = qx - cs.z, dy = qy - cs.w;
            x = dx * rot.x + dy * rot.y + cs.z;
            y = dx * rot.z + dy * rot.w + cs.w;
         }
         if (!(x >= bb.x && y >= bb.y && x <= bb.z && y <= bb.w) || (!SIMPLE && qv == 0.0f)) continue;
         const int ix = (int)((x - bb.x) * cs.x);
         const int iy = (int)((y - bb.y) * cs.y);
         if (ix >= P.width || iy >= P.height) continue;
         /* staging entries are counted in 32 bits (f3b_frame_begin refuses more): one
          * multiply-add, and one widening multiply-add for the address */
         /* staging entries are counted in 32 bits (f3b_frame_begin refuses more); the bucket's
          * place in the staging copy is scrambled (f3b_internal.h: F3HistLayout) */
         const uint32_t lin = (uint32_t)iy * (uint32_t)P.lay.pitch + (uint32_t)ix;
         const uint32_t bidx = rep_off + f3b_stage_index(P.lay, lin);

         /* rect.c:465-498: palette lookup */
         const int cmap_size = (int)(hf >> 20);
         const float findex = qc * (float)cmap_size;
         int pi = (int)findex;
         const float4 *dmap = reinterpret_cast<const float4 *>(sf + (c_samp(ci).lay.z & 0xffffu));
         float4 col;
         if (SIMPLE || (hf & HF_PLOT_READY)) {
            /* step lookup, every xform fully visible, dmap truncated on the host */
            pi = pi < 0 ? 0 : (pi >= cmap_size - 1 ? cmap_size - 1 : pi);
            if (AGG && P.bin_regions) {
               /* a record carries the colour as four bytes: taken from the rgba8 copy of the
                * table when the host made one (whole numbers <= 255) */
               uint32_t rgba;
               if (hf & HF_PAL_PACKED) rgba = reinterpret_cast<const uint32_t *>(dmap + 256)[pi];
               else {
                  col = dmap[pi];
                  rgba = (uint32_t)col.x | ((uint32_t)col.y << 8) | ((uint32_t)col.z << 16) | ((uint32_t)col.w << 24);
               }
               bin_r = __umulhi((uint32_t)iy, P.bin_row_magic);
               bin_a = (((uint32_t)iy - bin_r * P.bin_region_rows) * (uint32_t)P.lay.pitch + (uint32_t)ix) | (255u << 24);
               bin_c = rgba;
               continue;
            }
            /* (the 4-byte copy of the table was tried here twice - fewer shared-memory wavefronts,
             * eight more instructions to turn bytes into floats: -5 % and -3.5 %,
             * profiles/r02_palette_rgba8_ab.txt) */
            col = dmap[pi];
            float cnt = 255.0f;
            if (AGG) {
               float4 tc = col;
               float tn = cnt;
               const bool mine = f3b_bump_aggregate(P, lin, (unsigned long long)bidx, &tc, &tn,
                                                    P.probe_iters ? P.probe_counts : nullptr);
               if (!mine) continue;
               col = tc;
               cnt = tn;
            }
            f3b_red_v4(P.hist_color + bidx, col.x, col.y, col.z, col.w);
            if (COUNT) f3b_red_f32(P.hist_count + bidx, cnt);
            if (REPORT && (((uint32_t)i ^ tid) & (F3B_SAMPLE_EVERY - 1)) == 0)
               atomicExch(&s_ring.entry[atomicAdd(&s_ring.head, 1u) % F3B_SAMPLE_RING], (uint32_t)bidx);
            continue;
         }
         if (hf & HF_PAL_LINEAR) {
            float frac;
            if (pi < 0) {
               pi = 0;
               frac = 0.0f;
            } else if (pi >= cmap_size - 1) {
               pi = cmap_size - 2;
               frac = 1.0f;
            } else
               frac = findex - (float)pi;
            const float4 c0 = dmap[pi], c1 = dmap[pi + 1];
            col.x = c0.x * (1.0f - frac) + c1.x * frac;
            col.y = c0.y * (1.0f - frac) + c1.y * frac;
            col.z = c0.z * (1.0f - frac) + c1.z * frac;
            col.w = c0.w * (1.0f - frac) + c1.w * frac;
         } else {
            pi = pi < 0 ? 0 : (pi >= cmap_size - 1 ? cmap_size - 1 : pi);
            col = dmap[pi];
         }
         float cnt = 255.0f;
         if (qv != 1.0f) {
            col.x *= qv;
            col.y *= qv;
            col.z *= qv;
            col.w *= qv;
            cnt *= qv;
         }
         if (P.floor_adds) {
            col.x = floorf(col.x);
            col.y = floorf(col.y);
            col.z = floorf(col.z);
            col.w = floorf(col.w);
            cnt = floorf(cnt);
         }
         if (AGG) {
            if (P.bin_regions) { /* floor_adds: whole numbers <= 255 */
               bin_r = __umulhi((uint32_t)iy, P.bin_row_magic);
               bin_a = (((uint32_t)iy - bin_r * P.bin_region_rows) * (uint32_t)P.lay.pitch + (uint32_t)ix) | ((uint32_t)cnt << 24);
               bin_c = (uint32_t)col.x | ((uint32_t)col.y << 8) | ((uint32_t)col.z << 16) | ((uint32_t)col.w << 24);
               continue;
            }
            float4 tc = col;
            float tn = cnt;
            const bool mine = f3b_bump_aggregate(P, lin, (unsigned long long)bidx, &tc, &tn,
                                                 P.probe_iters ? P.probe_counts : nullptr);
            if (!mine) continue;
            col = tc;
            cnt = tn;
         }
         f3b_red_v4(P.hist_color + bidx, col.x, col.y, col.z, col.w);
         if (COUNT) f3b_red_f32(P.hist_count + bidx, cnt);
         if (REPORT && (((uint32_t)i ^ tid) & (F3B_SAMPLE_EVERY - 1)) == 0)
            atomicExch(&s_ring.entry[atomicAdd(&s_ring.head, 1u) % F3B_SAMPLE_RING], (uint32_t)bidx);
         } while (0);
         if (binned) {
            __syncwarp();
            f3b_bin_commit(P, wq, bin_a, bin_c, bin_r, blk_next, blk_end);
         }
         if (wu && (((unsigned)(i + (int)job.fuse) & 3u) == 3u) && i + 1 < n_loop) {
            /* deal the walkers out anew (transpose of the warps x lanes grid) */
            float *xs = reinterpret_cast<float *>(smem + P.wu_smem_off);
            xs[tid] = px;
            xs[nct + tid] = py;
            xs[2 * nct + tid] = pc;
            f3b_chain_barrier(nct);
            const uint32_t from = (tid & 31u) * (nct >> 5) + (tid >> 5);
            px = xs[from];
            py = xs[nct + from];
            pc = xs[2 * nct + from];
            f3b_chain_barrier(nct);
         }
      }
   }
   if (AGG && P.bin_regions) f3b_bin_flush(P, wq, blk_next, blk_end);

   /* fic.badvals += badcount (rect.c:408) */
   for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
   if ((tid & 31u) == 0 && bad) atomicAdd(P.badvals, (unsigned long long)bad);
   if (P.sweep_on) { /* all walkers of the CTA are through: release its sweeper warp */
      f3b_chain_barrier(nct);
      if (tid == 0) atomicExch(&s_done, 1);
   }
}

/* One application of one xform to caller-supplied points (parity tests of the variation
 * maths against the reference's apply_xform). in: {x, y, color}; out: {x, y, color, vis, bad} */
__global__ void f3b_debug_apply_kernel(const uint32_t *blob, int xform_index, const float *in, int n,
                                       float *out, uint64_t seed) {
   extern __shared__ __align__(16) uint32_t smem[];
   const float *sf = reinterpret_cast<const float *>(smem);
   const uint32_t words = blob[F3B_H_WORDS];
   for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) smem[i] = blob[i];
   __syncthreads();
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const uint32_t *xftab = smem + smem[F3B_H_OFF_XFTAB];
   F3Rng rng;
   f3b_rng_seed(rng, seed, 0u, (uint32_t)i);
   float qx, qy, qc, qv;
   const bool bad = f3b_apply_xform(sf, xftab[xform_index], in[3 * i], in[3 * i + 1], in[3 * i + 2],
                                    qx, qy, qc, qv, rng);
   out[5 * i] = qx;
   out[5 * i + 1] = qy;
   out[5 * i + 2] = qc;
   out[5 * i + 3] = qv;
   out[5 * i + 4] = bad ? 1.0f : 0.0f;
}

/* flam3_iterate (flam3.c:232-299) for many independent chains at once: thread t is a chain
 * that starts at start[t] (or a random point of [-1,1]^2) with colour 0, discards `fuse`
 * iterations and stores n_per_chain samples {x, y, colour, vis_adjusted} - the final xform
 * applied to the stored sample but not fed back, a bad value retried up to 4 times - exactly
 * the loop the render kernel runs, without camera and buckets. out[i * n_chains + t]. For
 * the analysis helpers either side of the render (flam3_estimate_bounding_box,
 * flam3_colorhist): one chain is latency bound on a GPU, a few hundred are not. */
__global__ void __launch_bounds__(128) f3b_orbit_kernel(const uint32_t *blob, int n_chains, int n_per_chain,
                                                        int fuse, uint64_t seed, const float *start,
                                                        float4 *out, unsigned long long *badvals) {
   extern __shared__ __align__(16) uint32_t smem[];
   const float *sf = reinterpret_cast<const float *>(smem);
   const uint32_t words = blob[F3B_H_WORDS];
   for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) smem[i] = blob[i];
   __syncthreads();
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= n_chains) return;
   const uint32_t cstride = smem[F3B_H_CHOOSER_STRIDE], choff = smem[F3B_H_OFF_CHOOSER];
   const uint32_t xaos = smem[F3B_H_XAOS], xb = smem[F3B_H_XF_BASE], xs = smem[F3B_H_XF_STRIDE];
   const int fin = (int)smem[F3B_H_FINAL];
   const float final_opacity = sf[F3B_H_FINAL_OPACITY];
   F3Rng rng;
   f3b_rng_seed(rng, seed, 0x0b17u, (uint32_t)t);
   float px = start ? start[2 * t] : rng.u11();
   float py = start ? start[2 * t + 1] : rng.u11();
   float pc = 0.0f;
   uint32_t lastxf = 0;
   unsigned int bad = 0;
   for (int i = -fuse; i < n_per_chain; i++) {
      float qx, qy, qc, qv;
      uint32_t fn = 0;
      for (int consec = 0;;) {
         const uint32_t k = rng.next() & (F3B_CHOOSER_GRAIN - 1u);
         const uint32_t ro = choff + (xaos ? lastxf * cstride : 0u);
         fn = 0;
         for (uint32_t j = 0; j < cstride; j++) fn += k >= smem[ro + j] ? 1u : 0u;
         const bool isbad = f3b_apply_xform(sf, xb + fn * xs, px, py, pc, qx, qy, qc, qv, rng);
         if (!isbad) break;
         bad++;
         if (++consec >= 5) break;
         px = qx;
         py = qy;
         pc = qc;
      }
      lastxf = fn + 1;
      px = qx;
      py = qy;
      pc = qc;
      if (fin >= 0) {
         bool apply = final_opacity == 1.0f;
         if (!apply) apply = rng.u01() < final_opacity;
         if (apply) {
            float fx, fy, fc, fv;
            f3b_apply_xform(sf, xb + (uint32_t)fin * xs, px, py, pc, fx, fy, fc, fv, rng);
            qx = fx;
            qy = fy;
            qc = fc;
         }
      }
      if (i >= 0) out[(size_t)i * n_chains + t] = make_float4(qx, qy, qc, qv);
   }
   if (bad) atomicAdd(badvals, (unsigned long long)bad);
}
int f3b_launch_orbits(const uint32_t *blob, size_t blob_bytes, int n_chains, int n_per_chain, int fuse,
                      uint64_t seed, const float *start, float4 *out, unsigned long long *badvals,
                      cudaStream_t stream) {
   cudaFuncSetAttribute(f3b_orbit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                        (int)(F3B_MAX_BLOB_WORDS * 4));
   f3b_orbit_kernel<<<(n_chains + 127) / 128, 128, blob_bytes, stream>>>(blob, n_chains, n_per_chain, fuse, seed,
                                                                          start, out, badvals);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}

int f3b_launch_debug_apply(const uint32_t *blob, size_t blob_bytes, int xform_index, const float *in,
                           int n, float *out, uint64_t seed, cudaStream_t stream) {
   cudaFuncSetAttribute(f3b_debug_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                        (int)(F3B_MAX_BLOB_WORDS * 4));
   f3b_debug_apply_kernel<<<(n + 127) / 128, 128, blob_bytes, stream>>>(blob, xform_index, in, n, out,
                                                                         seed);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}

int f3b_iterate_const_capacity(void) { return F3B_PARAM_SAMPLES; }

int f3b_iterate_max_blocks_per_sm(size_t smem_bytes, bool simple) {
   return f3b_iterate_blocks_per_sm(F3B_BLOCK, smem_bytes, simple);
}

int f3b_iterate_blocks_per_sm(int group_size, size_t smem_bytes, bool simple) {
   int nb = 0;
   /* the blob, plus the per-warp region queues of the binned scatter (AGG kernels) */
   const int smax = (int)(F3B_MAX_BLOB_WORDS * 4) + 8 * (32 + F3B_BIN_MAX_REGIONS * 130) * 4;
   const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
   static bool attrs_set = false; /* per process and device 0..n alike: the attribute is per function */
   if (!attrs_set) attrs_set = true;
#define F3B_SET1(C, S, M, W)                                                    \
   cudaFuncSetAttribute(f3b_iterate_kernel<C, S, M, W, false>, at, smax);       \
   cudaFuncSetAttribute(f3b_iterate_kernel<C, S, M, W, true>, at, smax)
#define F3B_SET(C, S)          \
   F3B_SET1(C, S, 0, false);   \
   F3B_SET1(C, S, 0, true);    \
   F3B_SET1(C, S, 1, false);   \
   F3B_SET1(C, S, 1, true);    \
   F3B_SET1(C, S, 2, false);   \
   F3B_SET1(C, S, 2, true)
   F3B_SET(true, 0);
   F3B_SET(false, 0);
   F3B_SET(true, 1);
   F3B_SET(false, 1);
   F3B_SET(true, 2);
   F3B_SET(false, 2);
#undef F3B_SET
#undef F3B_SET1
   /* group_size counts the walkers; the CTA has one more warp (the sweeper). The SIMPLE and the
    * generic instantiations have different register budgets (3 and 2 CTAs per SM). */
   cudaError_t e = simple ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f3b_iterate_kernel<true, 1, 1, true, false>,
                                                                          group_size + 32, smem_bytes)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f3b_iterate_kernel<true, 0, 1, true, false>,
                                                                          group_size + 32, smem_bytes);
   if (e != cudaSuccess) return 0;
   return nb;
}

int f3b_launch_iterate(F3IterateParams &p, const F3ConstSample *consts, int n_consts,
                       bool count_channel, bool simple, bool aggregate, int grid_blocks,
                       int group_size, size_t smem_bytes, cudaStream_t stream) {
   if (grid_blocks <= 0 || n_consts <= 0 || n_consts > F3B_PARAM_SAMPLES) return 0;
   if (group_size < 32 || group_size > F3B_BLOCK || (group_size & 31)) return 0;
   for (int i = 0; i < n_consts; i++) p.samp[i] = consts[i];
   p.chain_threads = (uint32_t)group_size;
   const int threads = group_size + (p.sweep_on ? 32 : 0);
   /* the sweeper drains the walkers' reports only where the walkers make them */
   const int mode = aggregate ? 1 : ((p.sweep_on && (p.sweep_mode & 8u)) ? 2 : 0);
   const uint32_t sweep_mode_asked = p.sweep_mode;
   if (mode != 2) p.sweep_mode &= ~8u;
#define F3B_LAUNCH_W(C, S, M, W)                                                                     \
   do {                                                                                              \
      if (n_consts == 1) f3b_iterate_kernel<C, S, M, W, true><<<grid_blocks, threads, smem_bytes, stream>>>(p); \
      else f3b_iterate_kernel<C, S, M, W, false><<<grid_blocks, threads, smem_bytes, stream>>>(p);  \
   } while (0)
#define F3B_LAUNCH_M(C, S, M)                    \
   do {                                          \
      if (p.warp_uniform) F3B_LAUNCH_W(C, S, M, true); \
      else F3B_LAUNCH_W(C, S, M, false);         \
   } while (0)
#define F3B_LAUNCH(C, S)                 \
   do {                                  \
      if (mode == 1) F3B_LAUNCH_M(C, S, 1);      \
      else if (mode == 2) F3B_LAUNCH_M(C, S, 2); \
      else F3B_LAUNCH_M(C, S, 0);                \
   } while (0)
   int kind = simple ? 2 : 0;
   for (int i = 0; i < n_consts && kind == 2; i++)
      if (consts[i].lay.x & F3B_HF_ANY_POST) kind = 1;
   if (count_channel) {
      if (kind == 2) F3B_LAUNCH(true, 2);
      else if (kind == 1) F3B_LAUNCH(true, 1);
      else F3B_LAUNCH(true, 0);
   } else {
      if (kind == 2) F3B_LAUNCH(false, 2);
      else if (kind == 1) F3B_LAUNCH(false, 1);
      else F3B_LAUNCH(false, 0);
   }
#undef F3B_LAUNCH_M
#undef F3B_LAUNCH_W
   p.sweep_mode = sweep_mode_asked;
#undef F3B_LAUNCH
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}
