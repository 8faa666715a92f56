/*
 * f3b_host.h — internal header of the flam3.h-facing host half (f3b_pack.c,
 * f3b_isaac.c, f3b_render.c). Needs the reference's PUBLIC header flam3.h on the include
 * path, exactly like flam3-render.c or any other libflam3 client.
 */
#ifndef F3B_HOST_H
#define F3B_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "flam3.h"

#include "../../include/flam3_b200.h"
#include "f3b_filters.h"
#include "f3b_pack.h"

/* buffer geometry of a frame (rect.c:611-686) */
typedef struct {
   int image_width, image_height; /* output pixels; height halved for fields */
   int out_width;                 /* output row pitch in pixels (doubled for fields) */
   int oversample;
   int filter_width, gutter_width, de_offset;
   int width, height;             /* bucket grid = fic.width, fic.height */
   int field;
   double pixel_aspect_ratio;
} f3b_frame_geom;

/* camera of one temporal sample (rect.c:785-827) */
typedef struct {
   double bounds[4];
   double rot[2][2];
   double ws0, hs1;
   double ppux, ppuy;
   double sample_density; /* cp.sample_density * 2^(2 zoom) */
} f3b_camera;

int f3b_camera_setup(const flam3_genome *cp, const f3b_frame_geom *g, f3b_camera *cam);
int f3b_chaos_enabled(const flam3_genome *cp, int nstd);
/* returns 0 ok, 1 genome error (message printed), 2 out of space */
int f3b_pack_sample(const flam3_genome *cp, const f3b_camera *cam, double color_scalar,
                    int cmap_size, int floor_adds, uint32_t *words, size_t cap_words,
                    size_t *used_words, int *count_is_alpha);

/* test hook: first index i with r_i >= thr[k] of the reference's running sum r += dr, by the
 * packer's closed form (closed_form=1) or by the plain loop (0) */
int f3b_debug_chooser_counts(double dr, unsigned n, const double *thr, int nthr, uint32_t *first,
                             int closed_form);

/* ISAAC as the reference runs it on LP64 (isaac.c + isaac.h): next 32-bit draw from ctx. */
unsigned long f3b_isaac_next(randctx *ctx);
void f3b_isaac_refill(randctx *ctx);
void f3b_isaac_skip(randctx *ctx, unsigned long n);

#endif
