/*
 * oracle/ref_wrap.c — TEST INFRASTRUCTURE (oracle). Compiles the UNMODIFIED reference
 * flam3.c (which textually includes rect.c six times) from where it lies under
 * $(F3_REF) (default /root/reference), adding capture hooks without touching the
 * reference sources: malloc/free inside this one translation unit are routed through
 * wrappers so that the render block allocated at rect.c:685-701
 *     [ buckets | accumulators | per-thread point storage ]
 * can be copied out just before rect.c:1257 frees it. At that moment `buckets` holds the
 * final histogram of the last batch and `accumulate` the post-density-estimation (and,
 * with earlyclip, post-tone-map) accumulator — exactly the intermediate states the
 * post-pass parity tests need.
 *
 * Nothing here is used by the product path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int    f3ref_capture_on = 0;
static void  *f3ref_big_ptr = NULL;
static size_t f3ref_big_size = 0;
static void  *f3ref_copy = NULL;
static size_t f3ref_copy_size = 0;
/* only blocks at least this large are candidates for "the render block" */
#define F3REF_MIN_BLOCK (64u * 1024u)

static void *f3ref_malloc_hook(size_t n) {
   void *p = malloc(n);
   if (f3ref_capture_on && p && n >= F3REF_MIN_BLOCK && n > f3ref_big_size) {
      f3ref_big_ptr = p;
      f3ref_big_size = n;
   }
   return p;
}

static void f3ref_free_hook(void *p) {
   if (f3ref_capture_on && p && p == f3ref_big_ptr) {
      free(f3ref_copy);
      f3ref_copy = malloc(f3ref_big_size);
      if (f3ref_copy) {
         memcpy(f3ref_copy, p, f3ref_big_size);
         f3ref_copy_size = f3ref_big_size;
      }
      f3ref_big_ptr = NULL;
      f3ref_big_size = 0;
   }
   free(p);
}

/* Control surface used by tests/harness (ctypes). */
void f3ref_capture(int on) {
   f3ref_capture_on = on;
   f3ref_big_ptr = NULL;
   f3ref_big_size = 0;
   if (!on) {
      free(f3ref_copy);
      f3ref_copy = NULL;
      f3ref_copy_size = 0;
   }
}
const void *f3ref_captured_block(size_t *size) {
   if (size) *size = f3ref_copy_size;
   return f3ref_copy;
}

/* Pull in the system headers flam3.c needs BEFORE redefining malloc/free so their
 * prototypes are not rewritten; the include guards make the later includes no-ops. */
#include <ctype.h>
#include <errno.h>
#include <libgen.h>
#include <limits.h>
#include <locale.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <time.h>
#include <unistd.h>
#include <libxml/parser.h>

#define malloc(n) f3ref_malloc_hook(n)
#define free(p) f3ref_free_hook(p)

#include "flam3.c"
