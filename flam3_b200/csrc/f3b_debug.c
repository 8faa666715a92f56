/*
 * f3b_debug.c — introspection entry points of the drop-in used by the parity tests: they
 * expose the host packer so its chooser thresholds, packed parameters and flags can be
 * compared with what the reference computes for the same genome. Not used by flam3_render.
 */
#include <string.h>

#include "f3b_host.h"

/* Packs genome `idx` of a parsed array as one temporal sample (weight 1, frame geometry of
 * a plain non-interlaced render with the genome's own size). Returns f3b_pack_sample's
 * code; *used receives the blob size in words. */
int f3b_debug_pack_genome(void *cps, int idx, int floor_adds, uint32_t *out, size_t cap_words,
                          size_t *used) {
   flam3_genome *g = (flam3_genome *)cps + idx;
   f3b_frame_geom geo;
   f3b_camera cam;
   memset(&geo, 0, sizeof(geo));
   geo.pixel_aspect_ratio = 1.0;
   geo.oversample = g->spatial_oversample;
   geo.image_width = g->width;
   geo.image_height = g->height;
   geo.out_width = g->width;
   geo.width = g->width * g->spatial_oversample;
   geo.height = g->height * g->spatial_oversample;
   if (f3b_camera_setup(g, &geo, &cam)) return 1;
   int alpha_ok = 0;
   return f3b_pack_sample(g, &cam, 1.0, 256, floor_adds, out, cap_words, used, &alpha_ok);
}
