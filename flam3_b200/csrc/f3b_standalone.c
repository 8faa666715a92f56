/*
 * f3b_standalone.c — the handful of libflam3 entry points f3b_render.c calls besides the ones
 * this repository replaces, for programs that link NO reference object at all (tools/f3b-render):
 *
 *   flam3_interpolate   flam3.c:717-802   the front end's own f3b_interpolate (f3b_frontend.c):
 *                                         a copy for one key frame, aligned and blended genomes
 *                                         between several
 *   clear_cp            flam3.c:1270      reset a genome, free its xforms
 *   flam3_init_frame    flam3.c:2494      seed the frame's ISAAC context from $isaac_seed (or the
 *                                         clock), with irandinit exactly as the reference runs it
 *                                         on LP64 (isaac.c:61-110: mix() on unmasked 64-bit words)
 *
 * Not part of libflam3_b200.so or of the drop-in libflam3.so (which get these from the
 * reference objects they are linked with).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "f3b_host.h"
#include "../../include/flam3_b200_frontend.h"

void clear_cp(flam3_genome *cp, int default_flag) { f3b_genome_init(cp, default_flag == flam3_defaults_on); }

void flam3_interpolate(flam3_genome *cps, int ncps, double time, double stagger, flam3_genome *result) {
   if (f3b_interpolate(cps, ncps, time, stagger, result)) fprintf(stderr, "flam3_interpolate: out of memory\n");
}

static void mix8(ub4 *v) { /* isaac.c:48-59, deliberately not masked to 32 bits */
   v[0] ^= v[1] << 11; v[3] += v[0]; v[1] += v[2];
   v[1] ^= v[2] >> 2;  v[4] += v[1]; v[2] += v[3];
   v[2] ^= v[3] << 8;  v[5] += v[2]; v[3] += v[4];
   v[3] ^= v[4] >> 16; v[6] += v[3]; v[4] += v[5];
   v[4] ^= v[5] << 10; v[7] += v[4]; v[5] += v[6];
   v[5] ^= v[6] >> 4;  v[0] += v[5]; v[6] += v[7];
   v[6] ^= v[7] << 8;  v[1] += v[6]; v[7] += v[0];
   v[7] ^= v[0] >> 9;  v[2] += v[7]; v[0] += v[1];
}

void f3b_isaac_seed(randctx *c) { /* irandinit(c, 1): randrsl holds the seed material */
   ub4 v[8];
   c->randa = c->randb = c->randc = 0;
   for (int i = 0; i < 8; i++) v[i] = 0x9e3779b9ul;
   for (int i = 0; i < 4; i++) mix8(v);
   for (int pass = 0; pass < 2; pass++)
      for (int i = 0; i < RANDSIZ; i += 8) {
         const ub4 *src = pass ? c->randmem : c->randrsl;
         for (int k = 0; k < 8; k++) v[k] += src[i + k];
         mix8(v);
         for (int k = 0; k < 8; k++) c->randmem[i + k] = v[k];
      }
   f3b_isaac_refill(c);
   c->randcnt = RANDSIZ;
}

void flam3_init_frame(flam3_frame *f) {
   const char *seed = getenv("isaac_seed");
   memset(f->rc.randrsl, 0, RANDSIZ * sizeof(ub4));
   if (!seed) {
      const long t = (long)time(0);
      for (int i = 0; i < RANDSIZ; i++) f->rc.randrsl[i] = (ub4)t;
   } else
      strncpy((char *)&f->rc.randrsl, seed, RANDSIZ * sizeof(ub4));
   f3b_isaac_seed(&f->rc);
}
