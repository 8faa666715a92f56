/*
 * f3b_variations.cuh — the 99 flam3 variations and apply_xform as fp32 device code.
 *
 * Behavioural specification: variations.c:142-1937 (the var functions),
 * variations.c:2129-2399 (apply_xform), flam3.c:2519-2544 (the 28-bit uniform draws).
 * This is a re-expression for one-walker-per-thread SIMT execution: xform records live in
 * shared memory as flat word arrays (f3b_pack.h), host-side precalcs are already folded
 * into the packed parameters, and all math is single precision. The per-variation
 * parameter order is the one documented in f3b_pack.c:pack_var_params.
 *
 * The variation switch is split in two so the hot loop stays small in registers and
 * instructions: the cheap, common variations (pure FMA / divide / sqrt / fast sincos) are
 * inlined in f3b_apply_xform; everything that needs a libm-class routine (pow, exp, log,
 * tan, sinh, fmod, atan2 inside the variation, ...) lives in the __noinline__
 * f3b_var_slow and costs a call only when a genome actually uses it.
 */
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "f3b_pack.h"

#define F3B_EPS 1e-10f           /* EPS, private.h:47 */
#define F3B_PI 3.14159265358979323846f
#define F3B_1_PI 0.31830988618379067154f
#define F3B_2_PI 0.63661977236758134308f
#define F3B_PI_2 1.57079632679489661923f
#define F3B_PI_4 0.78539816339744830962f

/* ------------------------------------------------------------------------------------
 * Per-walker random stream: xoshiro128++ (4 x 32-bit state, all-bits-good output, so
 * the `& 16383` xform pick (flam3.c:253) and the `& 1` coin flip (flam3.c:2541) are safe).
 * It stands in for the per-thread ISAAC context of rect.c:858-865; streams are derived
 * from the frame's ISAAC state by the host (f3b_render.c) and split per chain by hashing
 * (seed, temporal sample, chain) -> state, i.e. counter-based stream derivation.
 * ------------------------------------------------------------------------------------ */
struct F3Rng {
   uint32_t s0, s1, s2, s3;

   __device__ __forceinline__ uint32_t next() {
      uint32_t r = __funnelshift_l(s0 + s3, s0 + s3, 7) + s0;
      uint32_t t = s1 << 9;
      s2 ^= s0;
      s3 ^= s1;
      s1 ^= s2;
      s0 ^= s3;
      s2 ^= t;
      s3 = __funnelshift_l(s3, s3, 11);
      return r;
   }
   /* flam3_random_isaac_01: (irand & 0xfffffff) / 0xfffffff in [0,1] */
   __device__ __forceinline__ float u01() {
      return (float)(next() & 0xfffffffu) * (1.0f / 268435455.0f);
   }
   /* flam3_random_isaac_11: ((irand & 0xfffffff) - 0x7ffffff) / 0x7ffffff in [-1,1] */
   __device__ __forceinline__ float u11() {
      return (float)((int)(next() & 0xfffffffu) - 0x7ffffff) * (1.0f / 134217727.0f);
   }
   __device__ __forceinline__ uint32_t bit() { return next() & 1u; }
};

__host__ __device__ __forceinline__ uint64_t f3b_mix64(uint64_t z) {
   /* splitmix64 finaliser */
   z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
   z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
   return z ^ (z >> 31);
}

__device__ __forceinline__ void f3b_rng_seed(F3Rng &r, uint64_t seed, uint32_t sample_id,
                                             uint32_t chain) {
   uint64_t a = f3b_mix64(seed + 0x9e3779b97f4a7c15ULL * ((uint64_t)sample_id + 1));
   uint64_t b = f3b_mix64(a ^ (0xd1b54a32d192ed03ULL * ((uint64_t)chain + 1)));
   uint64_t c = f3b_mix64(b + 0x9e3779b97f4a7c15ULL);
   r.s0 = (uint32_t)b;
   r.s1 = (uint32_t)(b >> 32);
   r.s2 = (uint32_t)c;
   r.s3 = (uint32_t)(c >> 32) | 1u; /* never the all-zero state */
}

#define F3B_NOSPEC asm volatile("")

/* badvalue(), variations.c:22 */
__device__ __forceinline__ bool f3b_bad(float v) { return !(fabsf(v) <= 1e10f); }

/* pseudo-gaussian used by gaussian_blur / radial_blur / pre_blur: sum of 4 uniforms - 2 */
__device__ __forceinline__ float f3b_gauss4(F3Rng &rng) {
   float a = rng.u01();
   float b = rng.u01();
   float c = rng.u01();
   float d = rng.u01();
   return a + b + c + d - 2.0f;
}


/* sin/cos with a two-constant Cody-Waite reduction to [-pi, pi] and the SFU approximations
 * there (|abs err| ~ 1e-6). Arguments beyond 2^14 fall back to the exact libm path. */
__device__ __noinline__ void f3b_sincos_big(float x, float *s, float *c) { sincosf(x, s, c); }

__device__ __forceinline__ void f3b_sincos(float x, float *s, float *c) {
   if (fabsf(x) < 16384.0f) {
      const float k = rintf(x * 0.15915494309189535f);
      float r = fmaf(k, -6.28318548202514648f, x);
      r = fmaf(k, 1.74845553e-07f, r);
      *s = __sinf(r);
      *c = __cosf(r);
   } else
      f3b_sincos_big(x, s, c);
}
__device__ __forceinline__ float f3b_sin(float x) {
   float s, c;
   f3b_sincos(x, &s, &c);
   return s;
}
__device__ __forceinline__ float f3b_cos(float x) {
   float s, c;
   f3b_sincos(x, &s, &c);
   return c;
}


/* Fast transcendental helpers for the variation code. The chaos game is compared with the
 * reference statistically (image PSNR, visit frequencies), and per application to ~1e-5
 * relative (tests/test_gpu_variations.py); the SFU approximations (ex2/lg2/rcp, ~1e-6) are
 * well inside that, and a random genome spends most of its instructions in libm's slow
 * paths otherwise (profiles/r01_c5_*: powf 94, atan2f 44 per call site, sinh/cosh 20 each). */
__device__ __forceinline__ float f3b_pow(float x, float y) { /* x >= 0 */
   if (y == 0.0f) return 1.0f;
   return exp2f(y * __log2f(x));
}
__device__ __forceinline__ float f3b_exp(float x) { return __expf(x); }
__device__ __forceinline__ float f3b_log(float x) { return __logf(x); }
/* atan2 from a degree-15 odd least-squares polynomial on [0,1] (max error 4e-8 rad before
 * fp32 rounding) */
__device__ __forceinline__ float f3b_atan2(float y, float x) {
   const float ax = fabsf(x), ay = fabsf(y);
   const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
   if (mx == 0.0f) return 0.0f;
   const float a = __fdividef(mn, mx);
   const float s = a * a;
   float r = fmaf(s, -0.00407331f, 0.02194661f);
   r = fmaf(r, s, -0.05606318f);
   r = fmaf(r, s, 0.09656256f);
   r = fmaf(r, s, -0.13915802f);
   r = fmaf(r, s, 0.19948509f);
   r = fmaf(r, s, -0.33330107f);
   r = fmaf(r, s, 0.99999944f);
   r *= a;
   if (ay > ax) r = 1.57079632679f - r;
   if (x < 0.0f) r = 3.14159265359f - r;
   return copysignf(r, y);
}
/* sinh and cosh together from one exponential; the series keeps sinh accurate near 0 */
__device__ __forceinline__ void f3b_sinhcosh(float x, float *sh, float *ch) {
   const float ax = fabsf(x);
   const float e = __expf(ax), ei = __fdividef(1.0f, e);
   *ch = 0.5f * (e + ei);
   float s = 0.5f * (e - ei);
   if (ax < 0.25f) {
      const float x2 = ax * ax;
      s = ax * fmaf(x2, fmaf(x2, fmaf(x2, 1.98412698e-4f, 8.33333333e-3f), 0.16666667f), 1.0f);
   }
   *sh = copysignf(s, x);
}
__device__ __forceinline__ float f3b_sinh(float x) {
   float s, c;
   f3b_sinhcosh(x, &s, &c);
   return s;
}
__device__ __forceinline__ float f3b_cosh(float x) {
   float s, c;
   f3b_sinhcosh(x, &s, &c);
   return c;
}
__device__ __forceinline__ float f3b_tan(float x) {
   float s, c;
   f3b_sincos(x, &s, &c);
   return __fdividef(s, c);
}

/* per-point quantities shared by the variations (flam3_iter_helper, private.h:102-119) */
struct F3Pt {
   float tx, ty, sumsq, rad, atan_xy, atan_yx, sina, cosa;
};

/* variations of the out-of-line set that draw random numbers (bit i of word i/32) */
__device__ __forceinline__ bool f3b_var_is_random(uint32_t id) {
   const uint32_t m = id < 32u ? 0x80000000u : (id < 64u ? 0x093cb23fu : (id < 96u ? 0x00004000u : 0x00000000u));
   return (m >> (id & 31u)) & 1u;
}

/* The heavyweight variations, out of line. `p` points at the packed parameters in shared
 * memory. Two functions: the ones that draw random numbers carry the generator state in and
 * out (through local memory, the price of an out-of-line call); the deterministic majority
 * does not pay for that. */
__device__ __noinline__ float2 f3b_var_slow_det(uint32_t id, float w, const float *p, F3Pt pt) {
   const float tx = pt.tx, ty = pt.ty, sumsq = pt.sumsq, rad = pt.rad;
   const float atan_xy = pt.atan_xy, atan_yx = pt.atan_yx, sina = pt.sina, cosa = pt.cosa;
   float p0 = 0.0f, p1 = 0.0f;
   switch (id) {
      case 17: { /* popcorn: p = {c20, c21} */
         float dx = f3b_tan(3.0f * ty);
         float dy = f3b_tan(3.0f * tx);
         p0 += w * (tx + p[0] * f3b_sin(dx));
         p1 += w * (ty + p[1] * f3b_sin(dy));
      } break;
      case 18: { /* exponential */
         float dx = w * f3b_exp(tx - 1.0f), s, c;
         f3b_sincos(F3B_PI * ty, &s, &c);
         p0 += dx * c;
         p1 += dx * s;
      } break;
      case 19: { /* power */
         float r = w * f3b_pow(rad, sina);
         p0 += r * cosa;
         p1 += r * sina;
      } break;
      case 20: { /* cosine */
         float s, c;
         f3b_sincos(tx * F3B_PI, &s, &c);
         p0 += w * (c * f3b_cosh(ty));
         p1 -= w * (s * f3b_sinh(ty));
      } break;
      case 21: { /* rings: p = {dx} */
         float dx = p[0];
         float r = w * (fmodf(rad + dx, 2.0f * dx) - dx + rad * (1.0f - dx));
         p0 += r * cosa;
         p1 += r * sina;
      } break;
      case 22: { /* fan: p = {dx, dy, dx2} */
         float a = atan_xy, s, c;
         a += (fmodf(a + p[1], p[0]) > p[2]) ? -p[2] : p[2];
         f3b_sincos(a, &s, &c);
         float r = w * rad;
         p0 += r * c;
         p1 += r * s;
      } break;
      case 23: { /* blob: p = {low, high-low, waves} */
         float r = rad * (p[0] + p[1] * (0.5f + 0.5f * f3b_sin(p[2] * atan_xy)));
         p0 += w * sina * r;
         p1 += w * cosa * r;
      } break;
      case 24: { /* pdj: p = {a,b,c,d} */
         float nx1 = f3b_cos(p[1] * tx);
         float nx2 = f3b_sin(p[2] * tx);
         float ny1 = f3b_sin(p[0] * ty);
         float ny2 = f3b_cos(p[3] * ty);
         p0 += w * (ny1 - nx1);
         p1 += w * (nx2 - ny2);
      } break;
      case 25: { /* fan2: p = {dx, dy, dx2} */
         float a = atan_xy, s, c;
         float t = a + p[1] - p[0] * truncf((a + p[1]) / p[0]);
         a = (t > p[2]) ? a - p[2] : a + p[2];
         f3b_sincos(a, &s, &c);
         float r = w * rad;
         p0 += r * s;
         p1 += r * c;
      } break;
      case 38: { /* ngon: p = {power/2, b=2pi/sides, corners, circle} */
         float r_factor = f3b_pow(sumsq, p[0]);
         float b = p[1];
         float phi = atan_yx - b * floorf(atan_yx / b);
         if (phi > 0.5f * b) phi -= b;
         float amp = p[2] * (1.0f / (f3b_cos(phi) + F3B_EPS) - 1.0f) + p[3];
         amp /= (r_factor + F3B_EPS);
         p0 += w * tx * amp;
         p1 += w * ty * amp;
      } break;
      case 42: /* tangent */
         p0 += w * f3b_sin(tx) / f3b_cos(ty);
         p1 += w * f3b_tan(ty);
         break;
      case 46: { /* secant2 */
         float cr = f3b_cos(w * rad);
         float icr = 1.0f / cr;
         p0 += w * tx;
         p1 += (cr < 0.0f) ? w * (icr + 1.0f) : w * (icr - 1.0f);
      } break;
      case 49: { /* disc2: p = {timespi, sinadd, cosadd} */
         float s, c;
         f3b_sincos(p[0] * (tx + ty), &s, &c);
         float r = w * atan_xy * F3B_1_PI;
         p0 += (s + p[2]) * r;
         p1 += (c + p[1]) * r;
      } break;
      case 55: { /* bipolar: p = {ps = -pi/2 * shift} */
         float t = sumsq + 1.0f;
         float x2 = 2.0f * tx;
         float y = 0.5f * f3b_atan2(2.0f * ty, sumsq - 1.0f) + p[0];
         if (y > F3B_PI_2) y = -F3B_PI_2 + fmodf(y + F3B_PI_2, F3B_PI);
         else if (y < -F3B_PI_2) y = F3B_PI_2 - fmodf(F3B_PI_2 - y, F3B_PI);
         p0 += w * 0.25f * F3B_2_PI * f3b_log((t + x2) / (t - x2));
         p1 += w * F3B_2_PI * y;
      } break;
      case 58: { /* cell: p = {size, 1/size} */
         int x = (int)floorf(tx * p[1]);
         int y = (int)floorf(ty * p[1]);
         float dx = tx - (float)x * p[0];
         float dy = ty - (float)y * p[0];
         if (y >= 0) {
            if (x >= 0) { y *= 2; x *= 2; }
            else { y *= 2; x = -(2 * x + 1); }
         } else {
            if (x >= 0) { y = -(2 * y + 1); x *= 2; }
            else { y = -(2 * y + 1); x = -(2 * x + 1); }
         }
         p0 += w * (dx + (float)x * p[0]);
         p1 -= w * (dy + (float)y * p[0]);
      } break;
      case 60: { /* curve: p = {xamp, yamp, 1/xlen^2, 1/ylen^2} */
         p0 += w * (tx + p[0] * f3b_exp(-ty * ty * p[2]));
         p1 += w * (ty + p[1] * f3b_exp(-tx * tx * p[3]));
      } break;
      case 61: { /* edisc */
         float tmp = sumsq + 1.0f;
         float tmp2 = 2.0f * tx;
         float xmax = (sqrtf(tmp + tmp2) + sqrtf(tmp - tmp2)) * 0.5f;
         float a1 = f3b_log(xmax + sqrtf(xmax - 1.0f));
         float a2 = -acosf(tx / xmax);
         float ww = w / 11.57034632f, snv, csv;
         f3b_sincos(a1, &snv, &csv);
         float snhu = f3b_sinh(a2), cshu = f3b_cosh(a2);
         if (ty > 0.0f) snv = -snv;
         p0 += ww * cshu * csv;
         p1 += ww * snhu * snv;
      } break;
      case 62: { /* elliptic */
         float tmp = sumsq + 1.0f;
         float x2 = 2.0f * tx;
         float xmax = 0.5f * (sqrtf(tmp + x2) + sqrtf(tmp - x2));
         float a = tx / xmax;
         float b = 1.0f - a * a;
         float ssx = xmax - 1.0f;
         float ww = w / F3B_PI_2;
         b = b < 0.0f ? 0.0f : sqrtf(b);
         ssx = ssx < 0.0f ? 0.0f : sqrtf(ssx);
         p0 += ww * f3b_atan2(a, b);
         if (ty > 0.0f) p1 += ww * f3b_log(xmax + ssx);
         else p1 -= ww * f3b_log(xmax + ssx);
      } break;
      case 63: { /* escher: p = {vc, vd} */
         float lnr = 0.5f * f3b_log(sumsq), s, c;
         float m = w * f3b_exp(p[0] * lnr - p[1] * atan_yx);
         float n = p[0] * atan_yx + p[1] * lnr;
         f3b_sincos(n, &s, &c);
         p0 += m * c;
         p1 += m * s;
      } break;
      case 64: { /* foci */
         float expx = f3b_exp(tx) * 0.5f;
         float expnx = 0.25f / expx, s, c;
         f3b_sincos(ty, &s, &c);
         float t = w / (expx + expnx - c);
         p0 += t * (expx - expnx);
         p1 += t * s;
      } break;
      case 65: { /* lazysusan: p = {spin, space, twist, x, y} */
         float x = tx - p[3];
         float y = ty + p[4];
         float r = sqrtf(x * x + y * y);
         if (r < w) {
            float a = f3b_atan2(y, x) + p[0] + p[2] * (w - r), s, c;
            f3b_sincos(a, &s, &c);
            r = w * r;
            p0 += r * c + p[3];
            p1 += r * s - p[4];
         } else {
            r = w * (1.0f + p[1] / r);
            p0 += r * x + p[3];
            p1 += r * y - p[4];
         }
      } break;
      case 68: { /* modulus: p = {x, y} */
         float xr = 2.0f * p[0], yr = 2.0f * p[1];
         if (tx > p[0]) p0 += w * (-p[0] + fmodf(tx + p[0], xr));
         else if (tx < -p[0]) p0 += w * (p[0] - fmodf(p[0] - tx, xr));
         else p0 += w * tx;
         if (ty > p[1]) p1 += w * (-p[1] + fmodf(ty + p[1], yr));
         else if (ty < -p[1]) p1 += w * (p[1] - fmodf(p[1] - ty, yr));
         else p1 += w * ty;
      } break;
      case 69: { /* oscilloscope: p = {2*pi*freq, amplitude, separation, damping} */
         float t;
         if (p[3] == 0.0f) t = p[1] * f3b_cos(p[0] * tx) + p[2];
         else t = p[1] * f3b_exp(-fabsf(tx) * p[3]) * f3b_cos(p[0] * tx) + p[2];
         p0 += w * tx;
         if (fabsf(ty) <= t) p1 -= w * ty;
         else p1 += w * ty;
      } break;
      case 70: { /* polar2 */
         float p2v = w * F3B_1_PI;
         p0 += p2v * atan_xy;
         p1 += p2v * 0.5f * f3b_log(sumsq);
      } break;
      case 71: { /* popcorn2: p = {x, y, c} */
         p0 += w * (tx + p[0] * f3b_sin(f3b_tan(ty * p[2])));
         p1 += w * (ty + p[1] * f3b_sin(f3b_tan(tx * p[2])));
      } break;
      case 74: { /* split: p = {xsize*pi, ysize*pi} */
         if (f3b_cos(tx * p[0]) >= 0.0f) p1 += w * ty;
         else p1 -= w * ty;
         if (f3b_cos(ty * p[1]) >= 0.0f) p0 += w * tx;
         else p0 -= w * tx;
      } break;
      case 77: { /* wedge: p = {angle, hole, count, swirl, comp_fac} */
         float a = atan_yx + p[3] * rad, s, c;
         float cc = floorf((p[2] * a + F3B_PI) * F3B_1_PI * 0.5f);
         a = a * p[4] + cc * p[0];
         f3b_sincos(a, &s, &c);
         float r = w * (rad + p[1]);
         p0 += r * c;
         p1 += r * s;
      } break;
      case 79: { /* wedge_sph: p = {angle, hole, count, swirl, comp_fac} */
         float r = 1.0f / (rad + F3B_EPS), s, c;
         float a = atan_yx + p[3] * r;
         float cc = floorf((p[2] * a + F3B_PI) * F3B_1_PI * 0.5f);
         a = a * p[4] + cc * p[0];
         f3b_sincos(a, &s, &c);
         r = w * (r + p[1]);
         p0 += r * c;
         p1 += r * s;
      } break;
      case 80: { /* whorl: p = {inside, outside} */
         float a, s, c;
         if (rad < w) a = atan_yx + p[0] / (w - rad);
         else a = atan_yx + p[1] / (w - rad);
         f3b_sincos(a, &s, &c);
         p0 += w * rad * c;
         p1 += w * rad * s;
      } break;
      case 81: { /* waves2: p = {freqx, scalex, freqy, scaley} */
         p0 += w * (tx + p[1] * f3b_sin(ty * p[0]));
         p1 += w * (ty + p[3] * f3b_sin(tx * p[2]));
      } break;
      case 82: { /* exp */
         float e = f3b_exp(tx), s, c;
         f3b_sincos(ty, &s, &c);
         p0 += w * e * c;
         p1 += w * e * s;
      } break;
      case 83: /* log */
         p0 += w * 0.5f * f3b_log(sumsq);
         p1 += w * atan_yx;
         break;
      case 84: { /* sin */
         float s, c;
         f3b_sincos(tx, &s, &c);
         p0 += w * s * f3b_cosh(ty);
         p1 += w * c * f3b_sinh(ty);
      } break;
      case 85: { /* cos */
         float s, c;
         f3b_sincos(tx, &s, &c);
         p0 += w * c * f3b_cosh(ty);
         p1 -= w * s * f3b_sinh(ty);
      } break;
      case 86: { /* tan */
         float s, c;
         f3b_sincos(2.0f * tx, &s, &c);
         float den = 1.0f / (c + f3b_cosh(2.0f * ty));
         p0 += w * den * s;
         p1 += w * den * f3b_sinh(2.0f * ty);
      } break;
      case 87: { /* sec */
         float s, c;
         f3b_sincos(tx, &s, &c);
         float den = 2.0f / (f3b_cos(2.0f * tx) + f3b_cosh(2.0f * ty));
         p0 += w * den * c * f3b_cosh(ty);
         p1 += w * den * s * f3b_sinh(ty);
      } break;
      case 88: { /* csc */
         float s, c;
         f3b_sincos(tx, &s, &c);
         float den = 2.0f / (f3b_cosh(2.0f * ty) - f3b_cos(2.0f * tx));
         p0 += w * den * s * f3b_cosh(ty);
         p1 -= w * den * c * f3b_sinh(ty);
      } break;
      case 89: { /* cot */
         float s, c;
         f3b_sincos(2.0f * tx, &s, &c);
         float den = 1.0f / (f3b_cosh(2.0f * ty) - c);
         p0 += w * den * s;
         p1 -= w * den * f3b_sinh(2.0f * ty);
      } break;
      case 90: { /* sinh */
         float s, c;
         f3b_sincos(ty, &s, &c);
         p0 += w * f3b_sinh(tx) * c;
         p1 += w * f3b_cosh(tx) * s;
      } break;
      case 91: { /* cosh */
         float s, c;
         f3b_sincos(ty, &s, &c);
         p0 += w * f3b_cosh(tx) * c;
         p1 += w * f3b_sinh(tx) * s;
      } break;
      case 92: { /* tanh */
         float s, c;
         f3b_sincos(2.0f * ty, &s, &c);
         float den = 1.0f / (c + f3b_cosh(2.0f * tx));
         p0 += w * den * f3b_sinh(2.0f * tx);
         p1 += w * den * s;
      } break;
      case 93: { /* sech */
         float s, c;
         f3b_sincos(ty, &s, &c);
         float den = 2.0f / (f3b_cos(2.0f * ty) + f3b_cosh(2.0f * tx));
         p0 += w * den * c * f3b_cosh(tx);
         p1 -= w * den * s * f3b_sinh(tx);
      } break;
      case 94: { /* csch */
         float s, c;
         f3b_sincos(ty, &s, &c);
         float den = 2.0f / (f3b_cosh(2.0f * tx) - f3b_cos(2.0f * ty));
         p0 += w * den * f3b_sinh(tx) * c;
         p1 -= w * den * f3b_cosh(tx) * s;
      } break;
      case 95: { /* coth */
         float s, c;
         f3b_sincos(2.0f * ty, &s, &c);
         float den = 1.0f / (f3b_cosh(2.0f * tx) - c);
         p0 += w * den * f3b_sinh(2.0f * tx);
         p1 += w * den * s;
      } break;
      case 96: { /* auger: p = {sym, weight, freq, scale} */
         float s = f3b_sin(p[2] * tx);
         float t = f3b_sin(p[2] * ty);
         float dy = ty + p[1] * (p[3] * s * 0.5f + fabsf(ty) * s);
         float dx = tx + p[1] * (p[3] * t * 0.5f + fabsf(tx) * t);
         p0 += w * (tx + p[0] * (dx - tx));
         p1 += w * dy;
      } break;
      case 97: { /* flux: p = {2 + spread} */
         float xpw = tx + w, xmw = tx - w;
         float avgr = w * p[0] * sqrtf(sqrtf(ty * ty + xpw * xpw) / sqrtf(ty * ty + xmw * xmw));
         float avga = (f3b_atan2(ty, xmw) - f3b_atan2(ty, xpw)) * 0.5f, s, c;
         f3b_sincos(avga, &s, &c);
         p0 += avgr * c;
         p1 += avgr * s;
      } break;
   default:
      break;
   }
   return make_float2(p0, p1);
}

__device__ __noinline__ float2 f3b_var_slow_rnd(uint32_t id, float w, const float *p, F3Pt pt,
                                            F3Rng *rngp) {
   const float tx = pt.tx, ty = pt.ty, sumsq = pt.sumsq, rad = pt.rad;
   const float atan_xy = pt.atan_xy, atan_yx = pt.atan_yx, sina = pt.sina, cosa = pt.cosa;
   (void)atan_xy, (void)sina, (void)cosa; /* not every set of cases uses all of them */
   F3Rng rng = *rngp;
   float p0 = 0.0f, p1 = 0.0f;
   switch (id) {
      case 31: { /* noise */
         float s, c;
         f3b_sincos(rng.u01() * 2.0f * F3B_PI, &s, &c);
         float r = w * rng.u01();
         p0 += tx * r * c;
         p1 += ty * r * s;
      } break;
      case 32: { /* julian: p = {power, cn, rN} */
         float t_rnd = truncf(p[2] * rng.u01());
         float a = (atan_yx + 2.0f * F3B_PI * t_rnd) / p[0], s, c;
         float r = w * f3b_pow(sumsq, p[1]);
         f3b_sincos(a, &s, &c);
         p0 += r * c;
         p1 += r * s;
      } break;
      case 33: { /* juliascope: p = {power, cn, rN} */
         int t_rnd = (int)truncf(p[2] * rng.u01());
         float a, s, c;
         if ((t_rnd & 1) == 0) a = (2.0f * F3B_PI * (float)t_rnd + atan_yx) / p[0];
         else a = (2.0f * F3B_PI * (float)t_rnd - atan_yx) / p[0];
         f3b_sincos(a, &s, &c);
         float r = w * f3b_pow(sumsq, p[1]);
         p0 += r * c;
         p1 += r * s;
      } break;
      case 34: { /* blur */
         float s, c;
         f3b_sincos(rng.u01() * 2.0f * F3B_PI, &s, &c);
         float r = w * rng.u01();
         p0 += r * c;
         p1 += r * s;
      } break;
      case 35: { /* gaussian_blur */
         float s, c;
         f3b_sincos(rng.u01() * 2.0f * F3B_PI, &s, &c);
         float r = w * f3b_gauss4(rng);
         p0 += r * c;
         p1 += r * s;
      } break;
      case 36: { /* radial_blur: p = {spinvar, zoomvar} */
         float g = w * f3b_gauss4(rng), s, c;
         f3b_sincos(atan_yx + p[0] * g, &s, &c);
         float rz = p[1] * g - 1.0f;
         p0 += rad * c + rz * tx;
         p1 += rad * s + rz * ty;
      } break;
      case 37: { /* pie: p = {slices, rotation, thickness} */
         int sl = (int)(rng.u01() * p[0] + 0.5f);
         float a = p[1] + 2.0f * F3B_PI * ((float)sl + rng.u01() * p[2]) / p[0], s, c;
         float r = w * rng.u01();
         f3b_sincos(a, &s, &c);
         p0 += r * c;
         p1 += r * s;
      } break;
      case 41: { /* arch (weight used non-linearly, as in the reference) */
         float s, c;
         f3b_sincos(rng.u01() * w * F3B_PI, &s, &c);
         p0 += w * s;
         p1 += w * (s * s) / c;
      } break;
      case 44: { /* rays */
         float ang = w * rng.u01() * F3B_PI;
         float r = w / (sumsq + F3B_EPS);
         float tanr = w * f3b_tan(ang) * r;
         p0 += tanr * f3b_cos(tx);
         p1 += tanr * f3b_sin(ty);
      } break;
      case 45: { /* blade */
         float r = rng.u01() * w * rad, s, c;
         f3b_sincos(r, &s, &c);
         p0 += w * tx * (c + s);
         p1 += w * tx * (c - s);
      } break;
      case 47: { /* twintrian */
         float r = rng.u01() * w * rad, s, c;
         f3b_sincos(r, &s, &c);
         float diff = log10f(s * s) + c;
         if (f3b_bad(diff)) diff = -30.0f;
         p0 += w * tx * diff;
         p1 += w * tx * (diff - s * F3B_PI);
      } break;
      case 50: { /* super_shape: p = {pm_4, n2, n3, rnd, holes, pneg1_n1} */
         float st, ct;
         f3b_sincos(p[0] * atan_yx + F3B_PI_4, &st, &ct);
         float t1 = f3b_pow(fabsf(ct), p[1]);
         float t2 = f3b_pow(fabsf(st), p[2]);
         float r = w * ((p[3] * rng.u01() + (1.0f - p[3]) * rad) - p[4]) * f3b_pow(t1 + t2, p[5]) /
                   rad;
         p0 += r * tx;
         p1 += r * ty;
      } break;
      case 51: { /* flower: p = {petals, holes} */
         float r = w * (rng.u01() - p[1]) * f3b_cos(p[0] * atan_yx) / rad;
         p0 += r * tx;
         p1 += r * ty;
      } break;
      case 52: { /* conic: p = {eccentricity, holes} */
         float ct = tx / rad;
         float r = w * (rng.u01() - p[1]) * p[0] / (1.0f + p[0] * ct) / rad;
         p0 += r * tx;
         p1 += r * ty;
      } break;
      case 53: { /* parabola: p = {height, width} */
         float s, c;
         f3b_sincos(rad, &s, &c);
         float a = rng.u01();
         float b = rng.u01();
         p0 += p[0] * w * s * s * a;
         p1 += p[1] * w * c * b;
      } break;
      case 56: { /* boarders */
         float rx = rintf(tx), ry = rintf(ty);
         float ox = tx - rx, oy = ty - ry;
         if (rng.u01() >= 0.75f) {
            p0 += w * (ox * 0.5f + rx);
            p1 += w * (oy * 0.5f + ry);
         } else if (fabsf(ox) >= fabsf(oy)) {
            if (ox >= 0.0f) {
               p0 += w * (ox * 0.5f + rx + 0.25f);
               p1 += w * (oy * 0.5f + ry + 0.25f * oy / ox);
            } else {
               p0 += w * (ox * 0.5f + rx - 0.25f);
               p1 += w * (oy * 0.5f + ry - 0.25f * oy / ox);
            }
         } else {
            if (oy >= 0.0f) {
               p1 += w * (oy * 0.5f + ry + 0.25f);
               p0 += w * (ox * 0.5f + rx + ox / oy * 0.25f);
            } else {
               p1 += w * (oy * 0.5f + ry - 0.25f);
               p0 += w * (ox * 0.5f + rx - ox / oy * 0.25f);
            }
         }
      } break;
      case 59: { /* cpow: p = {va, vc, vd, power} */
         float lnr = 0.5f * f3b_log(sumsq), s, c;
         float ang = p[1] * atan_yx + p[2] * lnr + p[0] * floorf(p[3] * rng.u01());
         float m = w * f3b_exp(p[1] * lnr - p[2] * atan_yx);
         f3b_sincos(ang, &s, &c);
         p0 += m * c;
         p1 += m * s;
      } break;
      case 78: { /* wedge_julia: p = {cn, rN, power, count, cf, angle} */
         float r = w * f3b_pow(sumsq, p[0]), s, c;
         int t_rnd = (int)(p[1] * rng.u01());
         float a = (atan_yx + 2.0f * F3B_PI * (float)t_rnd) / p[2];
         float cc = floorf((p[3] * a + F3B_PI) * F3B_1_PI * 0.5f);
         a = a * p[4] + cc * p[5];
         f3b_sincos(a, &s, &c);
         p0 += r * c;
         p1 += r * s;
      } break;
   default:
      break;
   }
   *rngp = rng;
   return make_float2(p0, p1);
}

/*
 * Applies xform record `xf` (shared memory, 16-byte aligned, layout f3b_pack.h) to
 * (px,py,pc). Returns true when the result is a "bad value", in which case (qx,qy) are
 * fresh uniform draws (variations.c:2391-2395).
 */
/* Optional per-point precalcs (variations.c:2149-2170): pre_blur, sqrt, atan2s, sin/cos of
 * the angle. Out of line: the hot loop only pays a flag test when none is needed. */
__device__ __noinline__ void f3b_precalc(const float *sf, uint32_t xo, uint32_t flags, F3Pt *ptp,
                                         F3Rng *rngp) {
   F3Pt pt = *ptp;
   if (flags & F3B_XF_PREBLUR) { /* var67_pre_blur, variations.c:1480 */
      F3Rng rng = *rngp;
      const float g = sf[xo + F3B_X_PREBLUR] * f3b_gauss4(rng);
      float sa, ca;
      f3b_sincos(rng.u01() * 2.0f * F3B_PI, &sa, &ca);
      pt.tx += g * ca;
      pt.ty += g * sa;
      *rngp = rng;
   }
   if (flags & F3B_XF_NEED_SQRT) pt.rad = sqrtf(pt.tx * pt.tx + pt.ty * pt.ty);
   if (flags & F3B_XF_ATAN_XY) pt.atan_xy = f3b_atan2(pt.tx, pt.ty);
   if (flags & F3B_XF_ANGLES) {
      pt.sina = pt.tx / pt.rad;
      pt.cosa = pt.ty / pt.rad;
   }
   if (flags & F3B_XF_ATAN_YX) pt.atan_yx = f3b_atan2(pt.ty, pt.tx);
   *ptp = pt;
}

/* The variation records of one xform beyond the fixed-function slots:
 * {id | len << 8, weight, params[len]}, len even so each starts 8-byte aligned, stored at
 * word offsets [vi, vend) of the shared sample. Out of line for the same reason. */
__device__ __noinline__ float2 f3b_apply_varlist(const float *sf, uint32_t vi, uint32_t vend, F3Pt pt,
                                                 F3Rng *rngp) {
   F3Rng rng = *rngp;
   float tx = pt.tx, ty = pt.ty, sumsq = pt.sumsq, rad = pt.rad;
   float atan_xy = pt.atan_xy, sina = pt.sina, cosa = pt.cosa;
   float p0 = 0.0f, p1 = 0.0f;

   while (vi < vend) {
      /* Opaque redefinition of the per-point values: without it the compiler treats them as
       * invariant in this loop and hoists the arithmetic of EVERY case (range reductions,
       * reciprocals, ...) in front of it, to be executed for every point whatever
       * variations the xform uses (54 of 272 instructions per iteration, profiles/r01). */
      asm volatile("" : "+f"(tx), "+f"(ty), "+f"(sumsq), "+f"(rad), "+f"(atan_xy), "+f"(sina),
                        "+f"(cosa));
      const float2 hw = *reinterpret_cast<const float2 *>(sf + vi);
      const uint32_t w0 = __float_as_uint(hw.x);
      const uint32_t id = w0 & 0xffu;
      const float w = hw.y;
      const float *p = sf + vi + 2;
      vi += 2 + (w0 >> 8);
      /* F3B_NOSPEC keeps the compiler from if-converting the small cases into straight-line
       * code that would execute all of them for every point. */
      switch (id) {
      case 0: /* linear (normally in the fixed slot; kept for completeness) */
         F3B_NOSPEC;
         p0 = fmaf(w, tx, p0);
         p1 = fmaf(w, ty, p1);
         break;
      case 2: { /* spherical (idem) */
         F3B_NOSPEC;
         const float r2 = w / (sumsq + F3B_EPS);
         p0 = fmaf(r2, tx, p0);
         p1 = fmaf(r2, ty, p1);
      } break;
      case 1: /* sinusoidal */
         F3B_NOSPEC;
         p0 += w * f3b_sin(tx);
         p1 += w * f3b_sin(ty);
         break;
      case 3: { /* swirl */
         F3B_NOSPEC;
         float s, c;
         f3b_sincos(sumsq, &s, &c);
         p0 += w * (s * tx - c * ty);
         p1 += w * (c * tx + s * ty);
      } break;
      case 4: { /* horseshoe */
         F3B_NOSPEC;
         float r = w / (rad + F3B_EPS);
         p0 += (tx - ty) * (tx + ty) * r;
         p1 += 2.0f * tx * ty * r;
      } break;
      case 5: /* polar */
         F3B_NOSPEC;
         p0 += w * (atan_xy * F3B_1_PI);
         p1 += w * (rad - 1.0f);
         break;
      case 6: /* handkerchief */
         F3B_NOSPEC;
         p0 += w * rad * f3b_sin(atan_xy + rad);
         p1 += w * rad * f3b_cos(atan_xy - rad);
         break;
      case 7: { /* heart */
         F3B_NOSPEC;
         float s, c;
         f3b_sincos(rad * atan_xy, &s, &c);
         float r = w * rad;
         p0 += r * s;
         p1 -= r * c;
      } break;
      case 8: { /* disc */
         F3B_NOSPEC;
         float a = atan_xy * F3B_1_PI, s, c;
         f3b_sincos(F3B_PI * rad, &s, &c);
         p0 += w * s * a;
         p1 += w * c * a;
      } break;
      case 9: { /* spiral */
         F3B_NOSPEC;
         float r = rad + F3B_EPS, s, c;
         float r1 = w / r;
         f3b_sincos(r, &s, &c);
         p0 += r1 * (cosa + s);
         p1 += r1 * (sina - c);
      } break;
      case 10: { /* hyperbolic */
         F3B_NOSPEC;
         float r = rad + F3B_EPS;
         p0 += w * sina / r;
         p1 += w * cosa * r;
      } break;
      case 11: { /* diamond */
         F3B_NOSPEC;
         float s, c;
         f3b_sincos(rad, &s, &c);
         p0 += w * sina * c;
         p1 += w * cosa * s;
      } break;
      case 12: { /* ex */
         F3B_NOSPEC;
         float n0 = f3b_sin(atan_xy + rad);
         float n1 = f3b_cos(atan_xy - rad);
         float m0 = n0 * n0 * n0 * rad;
         float m1 = n1 * n1 * n1 * rad;
         p0 += w * (m0 + m1);
         p1 += w * (m0 - m1);
      } break;
      case 13: { /* julia */
         F3B_NOSPEC;
         float a = 0.5f * atan_xy, s, c;
         if (rng.bit()) a += F3B_PI;
         float r = w * sqrtf(rad);
         f3b_sincos(a, &s, &c);
         p0 += r * c;
         p1 += r * s;
      } break;
      case 14: { /* bent */
         F3B_NOSPEC;
         float nx = tx < 0.0f ? tx * 2.0f : tx;
         float ny = ty < 0.0f ? ty * 0.5f : ty;
         p0 += w * nx;
         p1 += w * ny;
      } break;
      case 15: { /* waves: p = {c10, dx2, c11, dy2} */
         F3B_NOSPEC;
         p0 += w * (tx + p[0] * f3b_sin(ty * p[1]));
         p1 += w * (ty + p[2] * f3b_sin(tx * p[3]));
      } break;
      case 16: { /* fisheye (x/y swap is intentional in the reference) */
         F3B_NOSPEC;
         float r = 2.0f * w / (rad + 1.0f);
         p0 += r * ty;
         p1 += r * tx;
      } break;
      case 26: { /* rings2: p = {dx} */
         F3B_NOSPEC;
         float dx = p[0];
         float r = rad;
         r += -2.0f * dx * truncf((r + dx) / (2.0f * dx)) + r * (1.0f - dx);
         p0 += w * sina * r;
         p1 += w * cosa * r;
      } break;
      case 27: { /* eyefish */
         F3B_NOSPEC;
         float r = (w * 2.0f) / (rad + 1.0f);
         p0 += r * tx;
         p1 += r * ty;
      } break;
      case 28: { /* bubble */
         F3B_NOSPEC;
         float r = w / (0.25f * sumsq + 1.0f);
         p0 += r * tx;
         p1 += r * ty;
      } break;
      case 29: /* cylinder */
         F3B_NOSPEC;
         p0 += w * f3b_sin(tx);
         p1 += w * ty;
         break;
      case 30: { /* perspective: p = {dist, vsin, vfcos} */
         F3B_NOSPEC;
         float t = 1.0f / (p[0] - ty * p[1]);
         p0 += w * p[0] * tx * t;
         p1 += w * p[2] * ty * t;
      } break;
      case 39: { /* curl: p = {c1, c2} */
         F3B_NOSPEC;
         float re = 1.0f + p[0] * tx + p[1] * (tx * tx - ty * ty);
         float im = p[0] * ty + 2.0f * p[1] * tx * ty;
         float r = w / (re * re + im * im);
         p0 += (tx * re + ty * im) * r;
         p1 += (ty * re - tx * im) * r;
      } break;
      case 40: { /* rectangles: p = {x, y} */
         F3B_NOSPEC;
         if (p[0] == 0.0f) p0 += w * tx;
         else p0 += w * ((2.0f * floorf(tx / p[0]) + 1.0f) * p[0] - tx);
         if (p[1] == 0.0f) p1 += w * ty;
         else p1 += w * ((2.0f * floorf(ty / p[1]) + 1.0f) * p[1] - ty);
      } break;
      case 43: { /* square */
         F3B_NOSPEC;
         float a = rng.u01();
         float b = rng.u01();
         p0 += w * (a - 0.5f);
         p1 += w * (b - 0.5f);
      } break;
      case 48: { /* cross */
         F3B_NOSPEC;
         float s = tx * tx - ty * ty;
         float r = w * sqrtf(1.0f / (s * s + F3B_EPS));
         p0 += tx * r;
         p1 += ty * r;
      } break;
      case 54: { /* bent2: p = {x, y} */
         F3B_NOSPEC;
         float nx = tx < 0.0f ? tx * p[0] : tx;
         float ny = ty < 0.0f ? ty * p[1] : ty;
         p0 += w * nx;
         p1 += w * ny;
      } break;
      case 57: { /* butterfly; 4/sqrt(3*pi) */
         F3B_NOSPEC;
         float wx = w * 1.3029400317411197908970256609023f;
         float y2 = ty * 2.0f;
         float r = wx * sqrtf(fabsf(ty * tx) / (F3B_EPS + tx * tx + y2 * y2));
         p0 += r * tx;
         p1 += r * y2;
      } break;
      case 66: { /* loonie */
         F3B_NOSPEC;
         float w2 = w * w;
         if (sumsq < w2) {
            float r = w * sqrtf(w2 / sumsq - 1.0f);
            p0 += r * tx;
            p1 += r * ty;
         } else {
            p0 += w * tx;
            p1 += w * ty;
         }
      } break;
      /* 67 pre_blur is applied before the loop and never packed as a loop entry */
      case 72: { /* scry */
         F3B_NOSPEC;
         float r = 1.0f / (rad * (sumsq + 1.0f / (w + F3B_EPS)));
         p0 += tx * r;
         p1 += ty * r;
      } break;
      case 73: { /* separation: p = {x^2, xinside, y^2, yinside} */
         F3B_NOSPEC;
         if (tx > 0.0f) p0 += w * (sqrtf(tx * tx + p[0]) - tx * p[1]);
         else p0 -= w * (sqrtf(tx * tx + p[0]) + tx * p[1]);
         if (ty > 0.0f) p1 += w * (sqrtf(ty * ty + p[2]) - ty * p[3]);
         else p1 -= w * (sqrtf(ty * ty + p[2]) + ty * p[3]);
      } break;
      case 75: { /* splits: p = {x, y} */
         F3B_NOSPEC;
         p0 += w * (tx >= 0.0f ? tx + p[0] : tx - p[0]);
         p1 += w * (ty >= 0.0f ? ty + p[1] : ty - p[1]);
      } break;
      case 76: { /* stripes: p = {1-space, warp} */
         F3B_NOSPEC;
         float roundx = floorf(tx + 0.5f);
         float offx = tx - roundx;
         p0 += w * (offx * p[0] + roundx);
         p1 += w * (ty + offx * offx * p[1]);
      } break;
      case 98: { /* mobius: p = {re_a, im_a, re_b, im_b, re_c, im_c, re_d, im_d} */
         F3B_NOSPEC;
         float re_u = p[0] * tx - p[1] * ty + p[2];
         float im_u = p[0] * ty + p[1] * tx + p[3];
         float re_v = p[4] * tx - p[5] * ty + p[6];
         float im_v = p[4] * ty + p[5] * tx + p[7];
         float rv = w / (re_v * re_v + im_v * im_v);
         p0 += rv * (re_u * re_v + im_u * im_v);
         p1 += rv * (im_u * re_v - re_u * im_v);
      } break;
      default: {
         float2 d;
         if (f3b_var_is_random(id)) {
            F3Rng tmp = rng; /* only this copy has its address taken */
            d = f3b_var_slow_rnd(id, w, p, pt, &tmp);
            rng = tmp;
         } else
            d = f3b_var_slow_det(id, w, p, pt);
         p0 += d.x;
         p1 += d.y;
      } break;
      }
   }
   *rngp = rng;
   return make_float2(p0, p1);
}

/*
 * apply_xform (variations.c:2129-2399). `sf` is the shared-memory copy of the packed
 * sample, `xo` the word offset of the xform record inside it (offsets, not pointers, keep
 * every access a 32-bit LDS address). Returns true when the result is a "bad value", in
 * which case (qx,qy) are fresh uniform draws (variations.c:2391-2395).
 *
 * The inlined part is what EVERY application needs: three LDS.128 of coefficients, the
 * affine map, the colour blend, and the two fixed-function variation slots (linear and
 * spherical, which nearly every flame uses). Everything else is reached through two flag
 * tests and runs out of line, so the hot loop's registers and instruction footprint are
 * those of the common case.
 */
__device__ __forceinline__ bool f3b_apply_xform(const float *__restrict__ sf, uint32_t xo, float px,
                                                float py, float pc, float &qx, float &qy, float &qc,
                                                float &qvis, F3Rng &rng, bool any_post = true) {
   const float4 *xf4 = reinterpret_cast<const float4 *>(sf + xo);
   const float4 a0 = xf4[0]; /* c00 c10 c20 c01 */
   const float4 a1 = xf4[1]; /* c11 c21 scolor 1-s */
   const float4 a2 = xf4[2]; /* vis, flags | list words << 8, linear weight, spherical weight */
   qc = fmaf(a1.w, pc, a1.z);
   qvis = a2.x;
   const uint32_t flags = __float_as_uint(a2.y);

   /* The common case needs tx, ty and their squares only; the point record with its angles and
    * radius is built on the two out-of-line paths that read it. */
   float tx = fmaf(a0.x, px, fmaf(a0.y, py, a0.z));
   float ty = fmaf(a0.w, px, fmaf(a1.x, py, a1.y));
   const uint32_t precalc = flags & (F3B_XF_PREBLUR | F3B_XF_NEED_SQRT | F3B_XF_ATAN_XY | F3B_XF_ANGLES | F3B_XF_ATAN_YX);
   F3Pt pt;
   if (precalc) {
      /* only these copies have their address taken; `rng` stays in registers */
      F3Rng tmp = rng;
      F3Pt tpt;
      tpt.tx = tx;
      tpt.ty = ty;
      tpt.rad = 0.0f;
      tpt.atan_xy = tpt.atan_yx = tpt.sina = tpt.cosa = 0.0f;
      tpt.sumsq = 0.0f;
      f3b_precalc(sf, xo, flags, &tpt, &tmp);
      rng = tmp;
      pt = tpt;
      tx = tpt.tx; /* pre_blur moves the point */
      ty = tpt.ty;
   }
   const float sumsq = tx * tx + ty * ty;

   /* fixed-function slots */
   float p0 = a2.z * tx, p1 = a2.z * ty;
   if (a2.w != 0.0f) {
      const float r2 = __fdividef(a2.w, sumsq + F3B_EPS);
      p0 = fmaf(r2, tx, p0);
      p1 = fmaf(r2, ty, p1);
   }
   /* the rest of the variation list, if any */
   if (flags >> 8) {
      if (!precalc) {
         pt.rad = 0.0f;
         pt.atan_xy = pt.atan_yx = pt.sina = pt.cosa = 0.0f;
      }
      pt.tx = tx;
      pt.ty = ty;
      pt.sumsq = sumsq;
      F3Rng tmp = rng;
      const float2 d = f3b_apply_varlist(sf, xo + F3B_X_VARS, xo + F3B_X_VARS + (flags >> 8), pt, &tmp);
      rng = tmp;
      p0 += d.x;
      p1 += d.y;
   }

   qx = p0;
   qy = p1;
   /* any_post is a compile-time constant in the kernels (false where no xform of the launch has
    * a post affine: most classic flames); if-converted, the block costs every iteration 8 issue
    * slots */
   if (any_post && (flags & F3B_XF_HAS_POST)) {
      const float4 b0 = xf4[3];                                              /* p00 p10 p20 p01 */
      const float2 b1 = *reinterpret_cast<const float2 *>(sf + xo + F3B_X_P11); /* p11 p21 */
      qx = fmaf(b0.x, p0, fmaf(b0.y, p1, b0.z));
      qy = fmaf(b0.w, p0, fmaf(b1.x, p1, b1.y));
   }
   if (f3b_bad(qx) || f3b_bad(qy)) {
      qx = rng.u11();
      qy = rng.u11();
      return true;
   }
   return false;
}
