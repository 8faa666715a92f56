/*
 * flam3_b200_dropin.h — the two entry points of libflam3 that the B200 build REPLACES, with
 * the reference's exact C ABI, plus the optional knobs the drop-in adds for embedding hosts.
 *
 * flam3_frame, flam3_genome, stat_struct and randctx are the reference's own types
 * (flam3.h:185-191, 194-523, 599-614; isaac.h:23-31). This header only forward-declares
 * them so it can be read without the reference tree; code that fills a flam3_frame includes
 * the reference's public header flam3.h, exactly like flam3-render.c does, and links
 * against dropin/_build/libflam3.so instead of the reference's libflam3.
 *
 * ABI facts the implementation relies on (x86-64 LP64, checked by tests/test_cpu_boundary.py):
 *   sizeof(randctx) = 288, sizeof(flam3_xform) = 3240, sizeof(flam3_genome) = 10728,
 *   sizeof(flam3_frame) = 360.
 */
#ifndef FLAM3_B200_DROPIN_H
#define FLAM3_B200_DROPIN_H

#include <stddef.h>
#include <stdint.h>

#include "flam3_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef flam3_included
struct flam3_frame_s;
typedef struct flam3_frame_opaque flam3_frame;
typedef struct stat_struct_opaque stat_struct;
#endif

/* Replaces flam3.c:3894-3932 (dispatch on bits / nthreads) and the six instantiations of
 * rect.c:render_rectangle behind it. Same arguments, same return convention (0 ok, 1 error
 * after a message on stderr), same side effects on *stats and on spec->rc; `out` is the
 * caller's nchan*bytes_per_channel*width*height buffer (flam3.h:621-623).
 *   bits 33 (default): float colour buckets with integer-valued adds + float accumulators
 *   bits 32: same buckets, integer accumulators, density estimation disabled
 *   bits 64: the adds are not truncated (the reference's double buckets)
 * nthreads is a hint: it only sets how many seed words are drawn from spec->rc. */
int flam3_render(flam3_frame *spec, void *out, int field, int nchan, int transp,
                 stat_struct *stats);

/* Replaces flam3.c:3874-3887; the same planning formula, so flam3-render's nstrips / use_mem
 * logic (flam3-render.c:36-101, 276-280) is unchanged. */
double flam3_render_memory_required(flam3_frame *spec);

/* ---- additions (not in the reference); all default to "off" ---------------------- */
/* The GPU context the drop-in renders with (created on first use on device $F3B_DEVICE). */
struct f3b_ctx *f3b_render_context(void);
/* Post-pass only: the next renders take this histogram (reference bits=33 layout,
 * uint32[5] per bucket, flam3.c:3824) instead of running the chaos game. NULL = off. */
void f3b_render_inject_histogram(const uint32_t *buckets5, size_t nbuckets);
/* Same, but the chaos game still runs and adds on top of the supplied buckets. */
void f3b_render_preload_histogram(const uint32_t *buckets5, size_t nbuckets);
/* Leave the finished image in device memory; `out` is not written. */
void f3b_render_set_device_only(int on);
/* This process runs chain groups g with g % nranks == rank (one process per GPU). */
void f3b_render_set_shard(int rank, int nranks);
/* Called after every batch's chaos game with the stream idle: sum the per-GPU histograms
 * here (f3b_hist_color_ptr / f3b_hist_count_ptr of include/flam3_b200.h). */
void f3b_render_set_exchange(void (*cb)(struct f3b_ctx *, void *), void *user);
/* Row-strip post-pass for nranks > 1: after the exchange every rank runs density estimation,
 * filtering and tone mapping only for image rows [rank*rows_per, (rank+1)*rows_per) with
 * rows_per = ceil(height/nranks), then calls `gather` with the stream idle. `image` is the
 * device image of nranks*rows_per rows of row_bytes bytes; gather must fill in the other
 * ranks' strips (an in-place all-gather of rows_per*row_bytes bytes per rank). NULL = off:
 * every rank post-processes the whole frame. */
/* By default both callbacks are called with the render stream idle and must have finished
 * their device work when they return. With stream_ordered = 1 they are called without any
 * host synchronisation and must enqueue their work on the render stream (f3b_stream_ptr of
 * include/flam3_b200.h), e.g. an NCCL collective issued on that stream. */
void f3b_render_set_stream_ordered(int on);
typedef void (*f3b_gather_fn)(struct f3b_ctx *, void *image, size_t row_bytes, int rows_per,
                              void *user);
void f3b_render_set_strip_gather(f3b_gather_fn cb, void *user);

/* ---- several GPUs on one frame (see include/flam3_b200.h) ---------------------------
 * In ONE process: set F3B_DEVICES ("0-7", "0,2", ...) before the first call; every
 * flam3_render then spreads its frame over those devices and `out` receives the whole image.
 * f3b_render_context_at(i) is the i-th of them.
 * In SEPARATE processes (one GPU each, e.g. under torchrun): f3b_render_set_shard(rank, n),
 * then tell every rank where all ranks' buckets and rank 0's image are (device pointers valid
 * in the calling process: CUDA IPC, f3b_ipc_export / f3b_ipc_open) and give a callback that
 * enqueues a barrier over all ranks on the render stream. Only rank 0's `out` is written.
 */
struct f3b_ctx *f3b_render_context_at(int i);
void f3b_render_set_peers(int nranks, const f3b_peer_hist *peers /* [nranks] */, void *image0);
void f3b_render_set_barrier(void (*cb)(struct f3b_ctx *, void *), void *user);

#ifdef __cplusplus
}
#endif
#endif
