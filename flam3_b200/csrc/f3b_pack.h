/*
 * f3b_pack.h — layout of a "packed temporal sample": everything the chaos-game kernel
 * needs for one interpolated control point, as a flat array of 32-bit words that is
 * copied verbatim into shared memory. Written by the host packer (f3b_pack.c, from a
 * flam3_genome) and read by the device (f3b_iterate.cu).
 *
 * It replaces what rect.c:754-884 hands each iter_thread: the thread's genome copy
 * (`fth[i].cp`), `fic.xform_distrib` (flam3.c:76-224), `fic.dmap` (rect.c:778-782) and
 * the camera constants (rect.c:785-827).
 *
 * All offsets are in 32-bit words from the start of the blob.
 */
#ifndef F3B_PACK_H
#define F3B_PACK_H

#include <stdint.h>

#define F3B_CHOOSER_GRAIN 16384u /* CHOOSE_XFORM_GRAIN, flam3.c:67 */
#define F3B_NVARS 99             /* flam3_nvariations, flam3.h */
#define F3B_MAX_BLOB_WORDS (48u * 1024u / 4u)

/* header word indices */
enum {
   F3B_H_WORDS = 0,     /* u32 total blob size in words */
   F3B_H_NSTD,          /* u32 number of standard (non-final) xforms */
   F3B_H_FINAL,         /* i32 index of the final xform in the xform table, or -1 if not enabled */
   F3B_H_XAOS,          /* u32 1 if the chooser has per-previous-xform rows (flam3.c:90-108) */
   F3B_H_FINAL_OPACITY, /* f32 opacity of the final xform (flam3.c:281-283) */
   F3B_H_PALMODE,       /* u32 0 = step, 1 = linear (rect.c:471-498) */
   F3B_H_CMAPSIZE,      /* u32 256, or 255 with apo_palette (rect.c:587) */
   F3B_H_ROTATE,        /* u32 camera rotation on (cp.rotate != 0, rect.c:419) */
   /* three float4 groups at words 8, 12, 16 (read with LDS.128 when a point is plotted) */
   F3B_H_ROT00, F3B_H_ROT01, F3B_H_ROT10, F3B_H_ROT11, /* f32 rect.c:815-818 */
   F3B_H_B0, F3B_H_B1, F3B_H_B2, F3B_H_B3,             /* f32 bounds, rect.c:809-812 */
   F3B_H_WS0, F3B_H_HS1,                               /* f32 buckets per unit, rect.c:819-822 */
   F3B_H_RCX, F3B_H_RCY,                               /* f32 rot_center */
   F3B_H_OFF_CHOOSER,   /* u32 -> u32 cum[rows][nstd]: entries of the 16384 table <= j */
   F3B_H_OFF_XFTAB,     /* u32 -> u32 word offset of each xform record [nxf] */
   F3B_H_OFF_PALETTE,   /* u32 -> float4 dmap[256] (16-byte aligned), then u32 rgba8[256] */
   F3B_H_NXF,           /* u32 number of xform records (nstd + 1 if final present) */
   F3B_H_PLOT,          /* u32 F3B_PLOT_*: what the bucket bump may skip */
   F3B_H_CHOOSER_STRIDE, /* u32 entries per chooser row: nstd rounded up to 4, pad = 16384 */
   F3B_H_XF_BASE,       /* u32 word offset of xform record 0 */
   F3B_H_XF_STRIDE,     /* u32 words between consecutive xform records (all padded alike) */
   F3B_H_SIZE = 32      /* header padded to 32 words */
};

/* F3B_H_PLOT bits. READY: every xform has vis_adjusted == 1, the palette lookup is "step"
 * and the dmap entries were already truncated on the host when the buckets are integer
 * (bits 32/33), so a hit adds dmap[i] as is (rect.c:499-505 with p[3]==1). */
#define F3B_PLOT_READY 1u
/* PACKED (only with READY): every dmap entry is a whole number in [0, 255], and a second copy of
 * the table follows the float4 one at F3B_H_OFF_PALETTE + 1024: 256 words r | g << 8 | b << 16 |
 * a << 24: what a record of the binned scatter carries. (For the direct path the 4-byte lookup
 * was measured and lost: fewer shared-memory wavefronts, but the eight instructions that turn
 * bytes into floats cost the issue-bound loop 5 %, profiles/r02_palette_rgba8_ab.txt.) */
#define F3B_PLOT_PACKED 2u

/* xform record: word indices relative to the record start (all f32 unless noted) */
/* Records start on a 16-byte boundary. The first 12 words are what every application of
 * the xform needs and are read with three LDS.128 (shared-memory bandwidth, 128 B/clk/SM
 * whatever the broadcast pattern, is what bounds the kernel: profiles/r01_v2); the post
 * affine and the pre_blur weight follow and are read only when flagged. */
enum {
   F3B_X_C00 = 0, F3B_X_C10, F3B_X_C20, /* tx = c00*x + c10*y + c20  (variations.c:2144) */
   F3B_X_C01, F3B_X_C11, F3B_X_C21,     /* ty = c01*x + c11*y + c21 */
   F3B_X_SCOLOR,   /* color_speed * color          (variations.c:2139) */
   F3B_X_ONEMS,    /* 1 - color_speed */
   F3B_X_VIS,      /* vis_adjusted                 (variations.c:2030, interpolation.c:22) */
   F3B_X_FLAGS,    /* u32 F3B_XF_* in bits 0-7 | (length in words of the variation list) << 8 */
   F3B_X_WLINEAR,  /* weight of var0 linear: fixed-function slot, not in the record list */
   F3B_X_WSPHERICAL, /* weight of var2 spherical: fixed-function slot */
   F3B_X_P00, F3B_X_P10, F3B_X_P20,     /* post affine (variations.c:2383-2385) */
   F3B_X_P01, F3B_X_P11, F3B_X_P21,
   F3B_X_PREBLUR,  /* pre_blur weight              (variations.c:2149-2150) */
   F3B_X_PAD0,
   F3B_X_VARS      /* first variation record: {u32 id | len<<8, f32 weight, f32 params[len]}, len =
                      nparams rounded up to even */
};

#define F3B_XF_HAS_POST 1u
#define F3B_XF_ATAN_XY 2u /* precalc_atan   = atan2(tx,ty)   variations.c:2158 */
#define F3B_XF_ANGLES 4u  /* precalc_sina/cosa = tx/r, ty/r  variations.c:2162 */
#define F3B_XF_ATAN_YX 8u /* precalc_atanyx = atan2(ty,tx)   variations.c:2168 */
#define F3B_XF_PREBLUR 16u
#define F3B_XF_NEED_SQRT 32u /* some active variation reads precalc_sqrt (variations.c:2154) */

/*
 * Number of packed parameter words per variation id (after id and weight). The order of
 * the parameters is fixed by f3b_pack.c:pack_var_params and consumed positionally by
 * f3b_variations.cuh. Host-side derived constants (the reference's *_precalc functions,
 * variations.c:1942-2002, and loop-invariant subexpressions) are folded in here in
 * double precision before narrowing to float.
 */
static const unsigned char f3b_var_nparams[F3B_NVARS] = {
   /*  0 linear       */ 0, /*  1 sinusoidal  */ 0, /*  2 spherical   */ 0, /*  3 swirl       */ 0,
   /*  4 horseshoe    */ 0, /*  5 polar       */ 0, /*  6 handkerch.  */ 0, /*  7 heart       */ 0,
   /*  8 disc         */ 0, /*  9 spiral      */ 0, /* 10 hyperbolic  */ 0, /* 11 diamond     */ 0,
   /* 12 ex           */ 0, /* 13 julia       */ 0, /* 14 bent        */ 0, /* 15 waves       */ 4,
   /* 16 fisheye      */ 0, /* 17 popcorn     */ 2, /* 18 exponential */ 0, /* 19 power       */ 0,
   /* 20 cosine       */ 0, /* 21 rings       */ 1, /* 22 fan         */ 3, /* 23 blob        */ 3,
   /* 24 pdj          */ 4, /* 25 fan2        */ 3, /* 26 rings2      */ 1, /* 27 eyefish     */ 0,
   /* 28 bubble       */ 0, /* 29 cylinder    */ 0, /* 30 perspective */ 3, /* 31 noise       */ 0,
   /* 32 julian       */ 3, /* 33 juliascope  */ 3, /* 34 blur        */ 0, /* 35 gaussian    */ 0,
   /* 36 radial_blur  */ 2, /* 37 pie         */ 3, /* 38 ngon        */ 4, /* 39 curl        */ 2,
   /* 40 rectangles   */ 2, /* 41 arch        */ 0, /* 42 tangent     */ 0, /* 43 square      */ 0,
   /* 44 rays         */ 0, /* 45 blade       */ 0, /* 46 secant2     */ 0, /* 47 twintrian   */ 0,
   /* 48 cross        */ 0, /* 49 disc2       */ 3, /* 50 super_shape */ 6, /* 51 flower      */ 2,
   /* 52 conic        */ 2, /* 53 parabola    */ 2, /* 54 bent2       */ 2, /* 55 bipolar     */ 1,
   /* 56 boarders     */ 0, /* 57 butterfly   */ 0, /* 58 cell        */ 2, /* 59 cpow        */ 4,
   /* 60 curve        */ 4, /* 61 edisc       */ 0, /* 62 elliptic    */ 0, /* 63 escher      */ 2,
   /* 64 foci         */ 0, /* 65 lazysusan   */ 5, /* 66 loonie      */ 0, /* 67 pre_blur    */ 0,
   /* 68 modulus      */ 2, /* 69 oscope      */ 4, /* 70 polar2      */ 0, /* 71 popcorn2    */ 3,
   /* 72 scry         */ 0, /* 73 separation  */ 4, /* 74 split       */ 2, /* 75 splits      */ 2,
   /* 76 stripes      */ 2, /* 77 wedge       */ 5, /* 78 wedge_julia */ 6, /* 79 wedge_sph   */ 5,
   /* 80 whorl        */ 2, /* 81 waves2      */ 4, /* 82 exp         */ 0, /* 83 log         */ 0,
   /* 84 sin          */ 0, /* 85 cos         */ 0, /* 86 tan         */ 0, /* 87 sec         */ 0,
   /* 88 csc          */ 0, /* 89 cot         */ 0, /* 90 sinh        */ 0, /* 91 cosh        */ 0,
   /* 92 tanh         */ 0, /* 93 sech        */ 0, /* 94 csch        */ 0, /* 95 coth        */ 0,
   /* 96 auger        */ 4, /* 97 flux        */ 1, /* 98 mobius      */ 8,
};

/* The per-temporal-sample job record (f3b_job) is part of the public ABI:
 * include/flam3_b200.h. */

#endif
