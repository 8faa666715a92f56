/*
 * f3b_isaac.c — the frame-level random source. flam3_frame carries an ISAAC context
 * (`rc`) whose evolution is part of the observable contract: render_rectangle draws
 * 16 words per worker thread per temporal sample from it (rect.c:858-865) and
 * flam3-render restores it between strips (flam3-render.c:320,336). The drop-in keeps
 * that contract by drawing the same words with the same generator and using them as seed
 * material for the GPU streams.
 *
 * Specification: isaac.c:25-46 (the round function) and isaac.h:49-52 (irand). RANDSIZL
 * is 4, so the state is 16 words, and `ub4` is `unsigned long` (isaacs.h:17): on LP64 the
 * state words are 64-bit containers in which the round function masks its results to 32
 * bits but memory words left by irandinit's unmasked mix() may carry high bits. The
 * arithmetic below therefore runs on 64-bit words with masks exactly where the reference
 * has them.
 */
#include "f3b_host.h"

#define N RANDSIZ
#define M32(x) ((x)&0xfffffffful)

void f3b_isaac_refill(randctx *c) {
   ub4 *mm = c->randmem, *rs = c->randrsl;
   ub4 a = c->randa;
   ub4 b = M32(c->randb + (++c->randc));
   for (int i = 0; i < N; i++) {
      const ub4 x = mm[i];
      ub4 mix;
      switch (i & 3) {
      case 0: mix = a << 13; break;
      case 1: mix = a >> 6; break;
      case 2: mix = a << 2; break;
      default: mix = a >> 16; break;
      }
      a = M32((a ^ mix) + mm[(i + N / 2) & (N - 1)]);
      const ub4 y = M32(mm[(x >> 2) & (N - 1)] + a + b);
      mm[i] = y;
      b = M32(mm[((y >> RANDSIZL) >> 2) & (N - 1)] + x);
      rs[i] = b;
   }
   c->randb = b;
   c->randa = a;
}

unsigned long f3b_isaac_next(randctx *c) {
   /* irand(): results are consumed from index 15 down to 0, then the pool is refilled */
   if (c->randcnt-- == 0) {
      f3b_isaac_refill(c);
      c->randcnt = N - 1;
   }
   return c->randrsl[c->randcnt];
}

/* n calls of f3b_isaac_next whose results nobody looks at: whole pools at a time */
void f3b_isaac_skip(randctx *c, unsigned long n) {
   while (n > 0) {
      if (c->randcnt == 0) {
         f3b_isaac_refill(c);
         c->randcnt = N;
      }
      const unsigned long take = n < (unsigned long)c->randcnt ? n : (unsigned long)c->randcnt;
      c->randcnt -= (ub4)take;
      n -= take;
   }
}
