/*
 * f3b_post.cu — the post-pass for sm_100a: log-density scaling + variable-width density
 * estimation, early-clip, spatial filter + tone map + quantisation.
 *
 * Replaces the reference's de_thread (rect.c:57-251), the no-DE accumulate loop
 * (rect.c:947-974), the early-clip sweep (rect.c:1086-1135) and the final filter loop
 * (rect.c:1137-1247) with palettes.c:199-349 (calc_alpha / calc_newrgb / hsv).
 *
 * Density estimation is reformulated from the reference's per-source SCATTER (racy across
 * thread swaths, rect.c:148-168) into a streaming classify pass plus a race-free per-target
 * GATHER over the tiles that wide kernels can reach; see the section below.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "f3b_internal.h"

#define DE_TILE 32       /* density-estimation tile edge, buckets */
#define DE_ROWS 8        /* thread rows per CTA: 32x8 threads, 4 buckets each */
#define DE_PER_THREAD (DE_TILE / DE_ROWS)
#define DE_NO_SOURCE 0xffffu

/* ---------------------------------------------------------------------------------- */
/* log scale without density estimation (rect.c:947-974)                               */
/* ---------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) f3b_logscale_kernel(const F3DensityParams P, size_t i0,
                                                           size_t n) {
   size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < n; i += stride) {
      const size_t y = i / (size_t)P.width, x = i - y * (size_t)P.width;
      const float4 c = f3b_view_color(P.hist, f3b_view_index(P.hist, (int)x, (int)y));
      if (c.w == 0.0f) continue;
      const float ls = (P.k1 * log1pf(c.w * P.k2)) / c.w;
      float4 a = P.accum[i];
      a.x += c.x * ls;
      a.y += c.y * ls;
      a.z += c.z * ls;
      a.w += c.w * ls;
      if (P.int_accum) { /* bits=32: unsigned accumulators truncate every add */
         a.x = floorf(a.x);
         a.y = floorf(a.y);
         a.z = floorf(a.z);
         a.w = floorf(a.w);
      }
      P.accum[i] = a;
   }
}

/* ---------------------------------------------------------------------------------- */
/* density estimation: classify (streaming) + gather (active tiles only)               */
/* ---------------------------------------------------------------------------------- */
/*
 * Pass 1, f3b_de_classify_kernel — one streaming sweep over the buckets. Every bucket gets
 * its kernel index (rect.c:84-121) written to a 2-byte plane; a bucket whose kernel has
 * radius 0 (the dense majority of a converged frame) contributes only to itself, so its
 * log-scaled colour times the centre coefficient goes straight into the accumulator here.
 * Each 32x32 tile records the largest radius >= 1 it holds.
 *
 * Pass 2, f3b_de_gather_kernel — one CTA per 32x32 tile; tiles that no wide kernel can reach
 * (the recorded radii of the tile and its neighbours are all 0) exit at once. The others
 * stage the sources with radius >= 1 of tile + halo in shared memory and every target sums
 * the kernels that reach it, in a fixed order, so the result is deterministic.
 *
 * Algorithmic traffic per bucket: 16 B histogram read + 16 B accumulator write (+16 B read
 * when this is not the first batch) + 2 B index plane.
 */
__device__ __forceinline__ float f3b_de_count(const F3DensityParams &P, int x, int y) {
   if (x < 0 || x >= P.width || y < 0 || y >= P.height) return 0.0f;
   if (y < P.hist.row0 || y >= P.hist.row0 + P.hist.rows) return 0.0f; /* outside the strip held */
   return f3b_view_count(P.hist, f3b_view_index(P.hist, x, y));
}

__global__ void __launch_bounds__(DE_TILE * DE_ROWS) f3b_de_classify_kernel(const F3DensityParams P) {
   const int os = P.oversample;
   const int ss = os / 2; /* floor(oversample/2), rect.c:61 */
   const double scfact = (os & 1) ? 1.0 : ((double)os / (os + 1.0)) * ((double)os / (os + 1.0));
   const int gx = blockIdx.x * DE_TILE + threadIdx.x;
   const int tile_y = P.class_tile_y0 + blockIdx.y;
   int my_maxrad = 0;
   float maxcnt = 0.0f;

   float4 c[DE_PER_THREAD];
   int gy[DE_PER_THREAD];
   bool live[DE_PER_THREAD];
#pragma unroll
   for (int q = 0; q < DE_PER_THREAD; q++) {
      gy[q] = tile_y * DE_TILE + threadIdx.y + q * DE_ROWS;
      live[q] = gx < P.width && gy[q] >= P.class_row_begin && gy[q] < P.class_row_end;
      c[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live[q]) c[q] = f3b_view_color(P.hist, f3b_view_index(P.hist, gx, gy[q]));
   }
#pragma unroll
   for (int q = 0; q < DE_PER_THREAD; q++) {
      if (!live[q]) continue;
      const size_t gi = (size_t)gx + (size_t)P.width * gy[q];
      const float cnt = P.hist.count ? f3b_view_count(P.hist, f3b_view_index(P.hist, gx, gy[q])) : c[q].w;
      maxcnt = fmaxf(maxcnt, cnt);
      unsigned kidx = DE_NO_SOURCE;
      float4 self = make_float4(0.f, 0.f, 0.f, 0.f);
      bool has_self = false;
      /* source range of the reference loops (rect.c:79-80), minus its out-of-range row */
      if (gx >= os - 1 && gx <= P.width - os && gy[q] >= os - 1 && gy[q] <= P.height - os &&
          cnt != 0.0f && c[q].w != 0.0f) {
         double f = (double)cnt;
         if (ss > 0) {
            f = 0.0;
            for (int jj = -ss; jj <= ss; jj++)
               for (int ii = -ss; ii <= ss; ii++) f += (double)f3b_de_count(P, gx + ii, gy[q] + jj);
         }
         f *= (1.0 / 255.0);
         f *= scfact;
         int k;
         if (f > (double)P.max_filtered_counts) k = P.max_filter_index;
         else if (f <= 100.0) k = (int)ceil(f) - 1;
         else {
            /* 100 + floor((f - 100)^curve) (rect.c:112): in float first; the double pow only
             * when the float power is too close to a whole number to be trusted (its relative
             * error is ~1e-6; the margin is 1e-4) */
            const float pf = exp2f(P.curve * log2f((float)(f - 100.0)));
            if (fabsf(pf - rintf(pf)) > 1e-4f * fmaxf(pf, 1.0f) + 1e-5f) k = 100 + (int)floorf(pf);
            else k = 100 + (int)floor(pow(f - 100.0, (double)P.curve));
         }
         if (k > P.max_filter_index) k = P.max_filter_index;
         if (k < 0) k = 0;
         kidx = (unsigned)k;
         const int rad = __ldg(P.de_radius + k);
         if (rad == 0) {
            const float ls = (P.k1 * log1pf(c[q].w * P.k2)) / c[q].w;
            const float coef = __ldg(P.de_coefs + (size_t)k * P.kernel_size) * ls;
            self = make_float4(coef * c[q].x, coef * c[q].y, coef * c[q].z, coef * c[q].w);
            has_self = true;
         } else {
            my_maxrad = max(my_maxrad, rad);
         }
      }
      P.de_kidx[gi] = (uint16_t)kidx;
      if (gy[q] >= P.row_begin && gy[q] < P.row_end) {
         if (P.overwrite_accum) {
            P.accum[gi] = self;
         } else if (has_self) {
            float4 a = P.accum[gi];
            a.x += self.x;
            a.y += self.y;
            a.z += self.z;
            a.w += self.w;
            P.accum[gi] = a;
         }
      }
   }
   /* largest wide-kernel radius of this tile, and the fullest bucket seen (as an integer:
    * the counts are whole numbers, or close enough for bits 64) */
   int my_maxcnt = (int)fminf(maxcnt, 2147483520.0f);
   for (int o = 16; o > 0; o >>= 1) {
      my_maxrad = max(my_maxrad, __shfl_xor_sync(0xffffffffu, my_maxrad, o));
      my_maxcnt = max(my_maxcnt, __shfl_xor_sync(0xffffffffu, my_maxcnt, o));
   }
   __shared__ int s_rad[DE_ROWS], s_cnt[DE_ROWS];
   if (threadIdx.x == 0) {
      s_rad[threadIdx.y] = my_maxrad;
      s_cnt[threadIdx.y] = my_maxcnt;
   }
   __syncthreads();
   if (threadIdx.x == 0 && threadIdx.y == 0) {
      int m = 0, c = 0;
      for (int i = 0; i < DE_ROWS; i++) {
         m = max(m, s_rad[i]);
         c = max(c, s_cnt[i]);
      }
      P.tile_maxrad[(size_t)tile_y * gridDim.x + blockIdx.x] = m;
      if (P.hot_max && c > 0) atomicMax(P.hot_max, (unsigned int)c);
   }
}

/* A kernel's coefficients are N_k * exp(-4.5 d^2 / h_k^2) inside the disc d <= h_k (Gaussian
 * with support 1.5 sampled at 1.5 d / h, filters.c:356-372), zero outside, and the reference
 * only visits |dx|,|dy| <= ceil(h_k) - 1 (rect.c:123). The staged gather evaluates that
 * instead of fetching the octant table: per staged source the colour already carries
 * ls * N_k, and a tap costs a compare against the largest d^2 inside the disc, one ex2 and
 * four FMAs - no dependent global load per tap, which is what held the table version back
 * (L1TEX 38 %, profiles/r01_v6_postpass). P.de_kparam[k] = {4.5 log2(e) / h_k^2, N_k}, P.de_d2max[k]
 * the largest d^2 the table holds a non-zero coefficient for. */
__global__ void __launch_bounds__(DE_TILE * DE_ROWS) f3b_de_gather_kernel(const F3DensityParams P) {
   extern __shared__ __align__(16) unsigned char de_smem[];
   __shared__ int s_reach;
   const int tiles_x = gridDim.x;
   const int tile_y = P.gather_tile_y0 + blockIdx.y;
   const int tid = threadIdx.y * DE_TILE + threadIdx.x;

   /* the widest kernel that can reach this tile */
   if (tid == 0) s_reach = 0;
   __syncthreads();
   {
      const int nt = (P.max_radius + DE_TILE - 1) / DE_TILE, span = 2 * nt + 1;
      int m = 0;
      for (int i = tid; i < span * span; i += DE_TILE * DE_ROWS) {
         const int tx = (int)blockIdx.x + (i % span) - nt, ty = tile_y + (i / span) - nt;
         if (tx >= 0 && tx < tiles_x && ty >= P.class_tile_y0 && ty < P.class_tile_y1)
            m = max(m, P.tile_maxrad[(size_t)ty * tiles_x + tx]);
      }
      if (m) atomicMax(&s_reach, m);
   }
   __syncthreads();
   const int R = s_reach;
   if (R == 0) return;

   const int SW = DE_TILE + 2 * R;
   float4 *s_col = reinterpret_cast<float4 *>(de_smem);                  /* colour * ls * N_k */
   float *s_c = reinterpret_cast<float *>(s_col + (size_t)SW * SW);       /* exponent scale */
   uint32_t *s_meta = reinterpret_cast<uint32_t *>(s_c + (size_t)SW * SW); /* d2max | radius << 16 */
   int *s_rowmax = reinterpret_cast<int *>(s_meta + (size_t)SW * SW);     /* widest kernel per staged row */
   const int x0 = blockIdx.x * DE_TILE - R, y0 = tile_y * DE_TILE - R;

   for (int i = tid; i < SW; i += DE_TILE * DE_ROWS) s_rowmax[i] = 0;
   __syncthreads();
   for (int i = tid; i < SW * SW; i += DE_TILE * DE_ROWS) {
      const int lx = i % SW, ly = i / SW;
      const int gx = x0 + lx, gy = y0 + ly;
      uint32_t meta = 0; /* radius 0 = nothing to gather from here */
      float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
      float cc = 0.f;
      if (gx >= 0 && gx < P.width && gy >= P.class_row_begin && gy < P.class_row_end) {
         const size_t gi = (size_t)gx + (size_t)P.width * gy;
         const unsigned k = P.de_kidx[gi];
         if (k != DE_NO_SOURCE) {
            const unsigned rad = __ldg(P.de_radius + k);
            if (rad) {
               const float4 c = f3b_view_color(P.hist, f3b_view_index(P.hist, gx, gy));
               const float2 kp = __ldg(P.de_kparam + k);
               const float ls = (P.k1 * log1pf(c.w * P.k2)) / c.w * kp.y;
               col = make_float4(c.x * ls, c.y * ls, c.z * ls, c.w * ls);
               cc = kp.x;
               meta = (uint32_t)__ldg(P.de_d2max + k) | (rad << 16);
               if ((int)rad > s_rowmax[ly]) atomicMax(&s_rowmax[ly], (int)rad);
            }
         }
      }
      s_col[i] = col;
      s_c[i] = cc;
      s_meta[i] = meta;
   }
   __syncthreads();

   const int gx = blockIdx.x * DE_TILE + threadIdx.x;
   if (gx >= P.width) return;
#pragma unroll 1
   for (int q = 0; q < DE_PER_THREAD; q++) {
      const int ly = threadIdx.y + q * DE_ROWS;
      const int gy = tile_y * DE_TILE + ly;
      if (gy < P.row_begin || gy >= P.row_end) continue;
      const int cx = threadIdx.x + R, cy = ly + R;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      bool any = false;
      for (int dy = -R; dy <= R; dy++) {
         const int ady = abs(dy);
         /* the whole warp shares cy: rows whose widest kernel is shorter than the vertical
          * distance are skipped, the others are scanned only as far as that kernel reaches */
         const int rm = s_rowmax[cy + dy];
         if (rm < ady) continue;
         const int rowbase = (cy + dy) * SW + cx;
         const int dy2 = dy * dy;
         for (int dx = -rm; dx <= rm; dx++) {
            const int n = rowbase + dx;
            const uint32_t meta = s_meta[n];
            const int rn = (int)(meta >> 16);
            const int d2 = dx * dx + dy2;
            if (max(abs(dx), ady) > rn || d2 > (int)(meta & 0xffffu) || rn == 0) continue;
            const float w = exp2f(-s_c[n] * (float)d2);
            const float4 c = s_col[n];
            acc.x = fmaf(w, c.x, acc.x);
            acc.y = fmaf(w, c.y, acc.y);
            acc.z = fmaf(w, c.z, acc.z);
            acc.w = fmaf(w, c.w, acc.w);
            any = true;
         }
      }
      if (!any) continue;
      const size_t gi = (size_t)gx + (size_t)P.width * gy;
      float4 a = P.accum[gi];
      a.x += acc.x;
      a.y += acc.y;
      a.z += acc.z;
      a.w += acc.w;
      P.accum[gi] = a;
   }
}

/* Same gather without the shared-memory staging, for kernels too wide for it: a tile plus a
 * halo of R buckets on every side needs (32 + 2R)^2 * 20 B, which passes the 227 KB a CTA can
 * have at R = 38 (estimator 9 at oversample 5, 20 at oversample 2, ...). The sources are then
 * read where they lie (the index plane is 2 B per bucket and L2-resident; colours are fetched
 * only for sources that do reach the target). Same summation order as the staged kernel, so
 * the two give identical accumulators. */
__global__ void __launch_bounds__(DE_TILE * DE_ROWS) f3b_de_gather_global_kernel(const F3DensityParams P) {
   __shared__ int s_reach;
   const int tiles_x = gridDim.x;
   const int tile_y = P.gather_tile_y0 + blockIdx.y;
   const int tid = threadIdx.y * DE_TILE + threadIdx.x;
   if (tid == 0) s_reach = 0;
   __syncthreads();
   {
      const int nt = (P.max_radius + DE_TILE - 1) / DE_TILE, span = 2 * nt + 1;
      int m = 0;
      for (int i = tid; i < span * span; i += DE_TILE * DE_ROWS) {
         const int tx = (int)blockIdx.x + (i % span) - nt, ty = tile_y + (i / span) - nt;
         if (tx >= 0 && tx < tiles_x && ty >= P.class_tile_y0 && ty < P.class_tile_y1)
            m = max(m, P.tile_maxrad[(size_t)ty * tiles_x + tx]);
      }
      if (m) atomicMax(&s_reach, m);
   }
   __syncthreads();
   const int R = s_reach;
   if (R == 0) return;
   const int gx = blockIdx.x * DE_TILE + threadIdx.x;
   if (gx >= P.width) return;
#pragma unroll 1
   for (int q = 0; q < DE_PER_THREAD; q++) {
      const int gy = tile_y * DE_TILE + threadIdx.y + q * DE_ROWS;
      if (gy < P.row_begin || gy >= P.row_end) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      bool any = false;
      for (int dy = -R; dy <= R; dy++) {
         const int sy = gy + dy;
         if (sy < P.class_row_begin || sy >= P.class_row_end) continue;
         const int ady = abs(dy);
         for (int dx = -R; dx <= R; dx++) {
            const int sx = gx + dx;
            if (sx < 0 || sx >= P.width) continue;
            const unsigned k = P.de_kidx[(size_t)sx + (size_t)P.width * sy];
            if (k == DE_NO_SOURCE) continue;
            const int rn = (int)__ldg(P.de_radius + k);
            const int adx = abs(dx);
            const int hi = max(adx, ady), lo = min(adx, ady);
            if (hi > rn || rn == 0) continue;
            const float4 c = f3b_view_color(P.hist, f3b_view_index(P.hist, sx, sy));
            const float ls = (P.k1 * log1pf(c.w * P.k2)) / c.w;
            const float coef = __ldg(P.de_coefs + (size_t)k * P.kernel_size + (hi * (hi + 1)) / 2 + lo);
            acc.x += coef * (c.x * ls);
            acc.y += coef * (c.y * ls);
            acc.z += coef * (c.z * ls);
            acc.w += coef * (c.w * ls);
            any = true;
         }
      }
      if (!any) continue;
      const size_t gi = (size_t)gx + (size_t)P.width * gy;
      float4 a = P.accum[gi];
      a.x += acc.x;
      a.y += acc.y;
      a.z += acc.z;
      a.w += acc.w;
      P.accum[gi] = a;
   }
}

int f3b_launch_density(const F3DensityParams &p_in, cudaStream_t stream) {
   if (p_in.row_end <= p_in.row_begin) return 1;
   if (p_in.n_kernels <= 0) {
      const size_t i0 = (size_t)p_in.width * p_in.row_begin, n = (size_t)p_in.width * p_in.row_end;
      int blocks = (int)((n - i0 + 255) / 256);
      if (blocks > 148 * 16) blocks = 148 * 16;
      f3b_logscale_kernel<<<blocks, 256, 0, stream>>>(p_in, i0, n);
      return cudaGetLastError() == cudaSuccess ? 1 : 0;
   }
   F3DensityParams p = p_in;
   const int R = p.max_radius;
   /* kernel indices are needed R rows beyond the target rows (sources that reach in) */
   p.class_row_begin = p.row_begin - R < 0 ? 0 : p.row_begin - R;
   p.class_row_end = p.row_end + R > p.height ? p.height : p.row_end + R;
   p.class_tile_y0 = p.class_row_begin / DE_TILE;
   p.class_tile_y1 = (p.class_row_end + DE_TILE - 1) / DE_TILE;
   p.gather_tile_y0 = p.row_begin / DE_TILE;
   const int gather_tile_y1 = (p.row_end + DE_TILE - 1) / DE_TILE;
   const int tiles_x = (p.width + DE_TILE - 1) / DE_TILE;

   const size_t SW = DE_TILE + 2 * R;
   size_t smem = SW * SW * (sizeof(float4) + sizeof(float) + sizeof(uint32_t)) + SW * sizeof(int);
   smem = (smem + 15) & ~(size_t)15;
   /* the staged kernel packs d2max into 16 bits: R <= 181; shared memory stops it long before */
   const bool staged = smem <= 220 * 1024 && !p.de_table_only;
   if (staged && cudaFuncSetAttribute(f3b_de_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem) != cudaSuccess)
      return 0;
   dim3 block(DE_TILE, DE_ROWS);
   f3b_de_classify_kernel<<<dim3(tiles_x, p.class_tile_y1 - p.class_tile_y0), block, 0, stream>>>(p);
   if (cudaGetLastError() != cudaSuccess) return 0;
   if (staged)
      f3b_de_gather_kernel<<<dim3(tiles_x, gather_tile_y1 - p.gather_tile_y0), block, smem, stream>>>(p);
   else
      f3b_de_gather_global_kernel<<<dim3(tiles_x, gather_tile_y1 - p.gather_tile_y0), block, 0, stream>>>(p);
   return cudaGetLastError() == cudaSuccess ? 2 : 0;
}

/* ---------------------------------------------------------------------------------- */
/* tone map helpers (palettes.c:199-349)                                               */
/* ---------------------------------------------------------------------------------- */
/* Written once for T = double and T = float. The image is defined by the double version (the
 * reference computes this pass in double and truncates to 8 or 16 bits; the <= 1 LSB gate is
 * met with < 0.1 % of the bytes differing). The float version is a FILTER in front of it: it
 * is ~10x cheaper (the double version keeps the FP64 pipe busy), and a pixel all of whose
 * channels land further than F3B_FIXUP_EPS from a truncation boundary gets the same bytes
 * from both; the others (a few %) are listed and redone in double by a second kernel. */
#define F3B_FIXUP_EPS 0.01f
template <typename T> __device__ __forceinline__ T f3b_pow(T a, T b);
template <> __device__ __forceinline__ double f3b_pow<double>(double a, double b) { return pow(a, b); }
/* the float pass is a filter in front of the double one (below): ex2(b lg2 a) is ~1e-6
 * accurate in relative terms, far inside the 0.01 it is trusted to */
template <> __device__ __forceinline__ float f3b_pow<float>(float a, float b) { return __powf(a, b); }

template <typename T> __device__ __forceinline__ T f3b_calc_alpha(T dens, T g, T linrange) {
   if (!(dens > T(0))) return T(0);
   if (dens < linrange) {
      const T funcval = f3b_pow<T>(linrange, g);
      const T frac = dens / linrange;
      return (T(1) - frac) * dens * (funcval / linrange) + frac * f3b_pow<T>(dens, g);
   }
   return f3b_pow<T>(dens, g);
}

template <typename T> __device__ void f3b_rgb2hsv(const T *rgb, T *hsv) {
   const T r = rgb[0], g = rgb[1], b = rgb[2];
   const T mx = fmax(r, fmax(g, b)), mn = fmin(r, fmin(g, b));
   const T del = mx - mn;
   const T s = (mx != T(0)) ? del / mx : T(0);
   T h = T(0);
   if (s != T(0)) {
      const T rc = (mx - r) / del, gc = (mx - g) / del, bc = (mx - b) / del;
      if (r == mx) h = bc - gc;
      else if (g == mx) h = T(2) + rc - bc;
      else if (b == mx) h = T(4) + gc - rc;
      if (h < T(0)) h += T(6);
   }
   hsv[0] = h;
   hsv[1] = s;
   hsv[2] = mx;
}

template <typename T> __device__ void f3b_hsv2rgb(const T *hsv, T *rgb) {
   T h = hsv[0];
   const T s = hsv[1], v = hsv[2];
   while (h >= T(6)) h -= T(6);
   while (h < T(0)) h += T(6);
   const int j = (int)floor(h);
   const T f = h - j;
   const T p = v * (T(1) - s), q = v * (T(1) - s * f), t = v * (T(1) - s * (T(1) - f));
   T r, g, b;
   switch (j) {
   case 1: r = q; g = v; b = p; break;
   case 2: r = p; g = v; b = t; break;
   case 3: r = p; g = q; b = v; break;
   case 4: r = t; g = p; b = v; break;
   case 5: r = v; g = p; b = q; break;
   default: r = v; g = t; b = p; break;
   }
   rgb[0] = r;
   rgb[1] = g;
   rgb[2] = b;
}

/* returns true when the hue-preserving clip (hsv round trip) was taken: the float filter does
 * not trust itself there */
template <typename T> __device__ bool f3b_calc_newrgb(const T *cbuf, T ls, T highpow, T *newrgb) {
   if (ls == T(0) || (cbuf[0] == T(0) && cbuf[1] == T(0) && cbuf[2] == T(0))) {
      newrgb[0] = newrgb[1] = newrgb[2] = T(0);
      return false;
   }
   /* most saturated channel */
   T maxa = T(-1), maxc = T(0);
   for (int i = 0; i < 3; i++) {
      const T a = ls * (cbuf[i] / T(255));
      if (a > maxa) {
         maxa = a;
         maxc = cbuf[i] / T(255);
      }
   }
   const T newls = T(255) / maxc;
   if (maxa > T(255) && highpow >= T(0)) {
      /* clip to full scale and desaturate instead of shifting hue */
      const T lsratio = f3b_pow<T>(newls / ls, highpow);
      T hsv[3];
      for (int i = 0; i < 3; i++) newrgb[i] = newls * (cbuf[i] / T(255)) / T(255);
      f3b_rgb2hsv<T>(newrgb, hsv);
      hsv[1] *= lsratio;
      f3b_hsv2rgb<T>(hsv, newrgb);
      for (int i = 0; i < 3; i++) newrgb[i] *= T(255);
      return true;
   }
   T adj = -highpow;
   if (adj > T(1)) adj = T(1);
   if (maxa <= T(255)) adj = T(1);
   for (int i = 0; i < 3; i++) newrgb[i] = ((T(1) - adj) * newls + adj * ls) * (cbuf[i] / T(255));
   return false;
}

/* gamma / vibrancy / highlight / background for one 4-vector t (rect.c:1100-1130 and
 * rect.c:1172-1210). On return t[0..2] are clamped to [0,255] and t[3] is alpha. */
template <typename T> __device__ __forceinline__ bool f3b_tonemap(T *t, const F3FinalParams &P) {
   T alpha, ls;
   if (t[3] <= T(0)) {
      alpha = T(0);
      ls = T(0);
   } else {
      const T tmp = t[3] / T(255);
      alpha = f3b_calc_alpha<T>(tmp, (T)P.g, (T)P.linrange);
      ls = (T)P.vibrancy * T(256) * alpha / tmp;
      if (alpha < T(0)) alpha = T(0);
      if (alpha > T(1)) alpha = T(1);
   }
   T newrgb[3];
   const bool hard = f3b_calc_newrgb<T>(t, ls, (T)P.highpow, newrgb);
   /* the reference always evaluates (1 - vibrancy) * 256 * pow(t/255, g) (rect.c:1195); with
    * the default vibrancy of exactly 1 the product is an exact zero whenever t >= 0 (pow is
    * finite), and three of the four pows per pixel can be skipped; a negative t (filters
    * with negative lobes) keeps the reference's NaN */
   const T inv_vib = T(1) - (T)P.vibrancy;
   for (int i = 0; i < 3; i++) {
      T a = newrgb[i];
      if (inv_vib != T(0) || t[i] < T(0)) a += inv_vib * T(256) * f3b_pow<T>(t[i] / T(255), (T)P.g);
      if (P.nchan <= 3 || P.transp == 0) a += (T(1) - alpha) * (T)P.background[i];
      else a = (alpha > T(0)) ? a / alpha : T(0);
      if (a > T(255)) a = T(255);
      if (a < T(0)) a = T(0);
      t[i] = a;
   }
   t[3] = alpha;
   return hard;
}

__global__ void __launch_bounds__(256) f3b_earlyclip_kernel(const F3FinalParams P, size_t i0, size_t n) {
   size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < n; i += stride) {
      const float4 a = P.accum[i];
      double t[4] = {(double)a.x, (double)a.y, (double)a.z, (double)a.w};
      f3b_tonemap<double>(t, P);
      P.accum[i] = make_float4((float)t[0], (float)t[1], (float)t[2], (float)t[3]);
   }
}

/* one output pixel: filter_width^2 taps in the reference's order (rect.c:1146-1160, ii outer,
 * jj inner), tone map, quantise. Returns true (float version only) when some channel is too
 * close to a truncation boundary for the float result to be trusted. */
template <typename T, typename FT>
__device__ __forceinline__ bool f3b_final_pixel(const F3FinalParams &P, int i, int j, const float4 *src,
                                                unsigned spitch, const FT *filt) {
   /* src: the accumulator value under the pixel's first tap, rows spitch apart (the
    * accumulator itself, or the CTA's tile of it in shared memory); filt: the filter */
   const int fw = P.filter_width;
   T t[4] = {T(0), T(0), T(0), T(0)};
   for (int ii = 0; ii < fw; ii++) {
      for (int jj = 0; jj < fw; jj++) {
         const T k = (T)filt[ii + jj * fw];
         const float4 a = src[(unsigned)ii + (unsigned)jj * spitch];
         t[0] += k * (T)a.x;
         t[1] += k * (T)a.y;
         t[2] += k * (T)a.z;
         t[3] += k * (T)a.w;
      }
   }
   bool unsure = false;
   if (sizeof(T) == 4 && t[0] == T(0) && t[1] == T(0) && t[2] == T(0) && t[3] == T(0)) {
      /* nothing under the filter (most of a sparse frame): the pixel is the background, which
       * sits ON a truncation boundary when it is a whole number (black). Exact zeros need no
       * filtering: take the double formula directly, it costs nothing here. */
      double td[4] = {0.0, 0.0, 0.0, 0.0};
      if (!P.earlyclip) f3b_tonemap<double>(td, P);
      unsigned char *o = reinterpret_cast<unsigned char *>(P.out) + ((size_t)i + (size_t)j * P.image_width) * P.nchan;
      for (int c = 0; c < 3; c++) o[c] = (unsigned char)fmin(fmax(td[c], 0.0), 255.0);
      if (P.nchan > 3) o[3] = P.transp ? (unsigned char)(fmin(fmax(td[3], 0.0), 1.0) * 255.0) : 255;
      return false;
   }
   if (!P.earlyclip) unsure = f3b_tonemap<T>(t, P);
   T alpha = t[3];
   if (alpha > T(1)) alpha = T(1);
   if (alpha < T(0)) alpha = T(0);
   const size_t pix = (size_t)i + (size_t)j * P.image_width;
   if (P.bytes_per_channel == 2) {
      unsigned short *o = reinterpret_cast<unsigned short *>(P.out) + pix * P.nchan;
      for (int c = 0; c < 3; c++) {
         T a = t[c];
         if (a > T(255)) a = T(255);
         if (a < T(0)) a = T(0);
         o[c] = (unsigned short)(a * T(256));
      }
      if (P.nchan > 3) o[3] = P.transp ? (unsigned short)(alpha * T(65535)) : 65535;
      return false; /* 16 bit is always rendered by the double version */
   }
   unsigned char *o = reinterpret_cast<unsigned char *>(P.out) + pix * P.nchan;
   for (int c = 0; c < 3; c++) {
      T a = t[c];
      if (!(a == a)) unsure = true; /* NaN: let the double version decide */
      if (a > T(255)) a = T(255);
      if (a < T(0)) a = T(0);
      /* distance to the nearest integer; values clamped at 0 or 255 sit ON a boundary unless
       * the unclamped value was clearly beyond it */
      if (sizeof(T) == 4) {
         const float v = (float)t[c];
         if (fabsf(v - rintf(v)) < F3B_FIXUP_EPS && v > -F3B_FIXUP_EPS && v < 255.0f + F3B_FIXUP_EPS) unsure = true;
      }
      o[c] = (unsigned char)a;
   }
   if (P.nchan > 3) {
      if (sizeof(T) == 4 && P.transp) {
         const float v = (float)alpha * 255.0f;
         if (fabsf(v - rintf(v)) < F3B_FIXUP_EPS) unsure = true;
      }
      o[3] = P.transp ? (unsigned char)(alpha * T(255)) : 255;
   }
   return unsure;
}

__device__ __forceinline__ const float4 *f3b_final_origin(const F3FinalParams &P, int i, int j) {
   return P.accum + (size_t)(P.de_offset + i * P.oversample) + (size_t)(P.de_offset + j * P.oversample) * P.width;
}
/* pass 1: every pixel in float; pixels the float result cannot be trusted for are listed. A CTA
 * renders 32x8 pixels from a shared-memory tile of the accumulator (each value is a tap of up
 * to filter_width^2 pixels: the tile turns 25-64 L2 loads per pixel into ~2) and a float copy
 * of the filter. tile_w == 0: no tile (it would not fit), taps come from global memory. */
__global__ void __launch_bounds__(256) f3b_final_kernel(const F3FinalParams P, int tile_w, int tile_h) {
   extern __shared__ __align__(16) unsigned char fin_smem[];
   const int i = blockIdx.x * blockDim.x + threadIdx.x; /* output column */
   const int j = P.out_row_begin + blockIdx.y * blockDim.y + threadIdx.y; /* output row */
   const bool inside = i < P.image_width && j < P.out_row_end;
   if (!P.fixup_list) { /* 16 bit output, or the filter switched off: all in double */
      if (inside) f3b_final_pixel<double, double>(P, i, j, f3b_final_origin(P, i, j), (unsigned)P.width, P.filter);
      return;
   }
   float *s_filt = reinterpret_cast<float *>(fin_smem);
   const int nf = P.filter_width * P.filter_width;
   float4 *s_tile = reinterpret_cast<float4 *>(fin_smem + (((size_t)nf * 4 + 15) & ~(size_t)15));
   const int tid = threadIdx.y * blockDim.x + threadIdx.x;
   for (int k = tid; k < nf; k += 256) s_filt[k] = (float)__ldg(P.filter + k);
   if (tile_w) {
      const int bx = P.de_offset + (int)(blockIdx.x * blockDim.x) * P.oversample;
      const int by = P.de_offset + (P.out_row_begin + (int)(blockIdx.y * blockDim.y)) * P.oversample;
      for (int k = tid; k < tile_w * tile_h; k += 256) {
         const int lx = k % tile_w, ly = k / tile_w;
         const int gx = bx + lx, gy = by + ly;
         s_tile[k] = (gx < P.width && gy < P.height) ? __ldg(P.accum + (size_t)gx + (size_t)gy * P.width)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
   }
   __shared__ unsigned int s_n, s_base;
   if (tid == 0) s_n = 0;
   __syncthreads();
   bool unsure = false;
   if (inside) {
      const float4 *src = tile_w ? s_tile + (size_t)threadIdx.x * P.oversample + (size_t)threadIdx.y * P.oversample * tile_w
                                 : f3b_final_origin(P, i, j);
      unsure = f3b_final_pixel<float, float>(P, i, j, src, tile_w ? (unsigned)tile_w : (unsigned)P.width, s_filt);
   }
   /* one global atomic per CTA, not per pixel: 10^5 atomics on one address cost more than the pass */
   unsigned int slot = 0;
   if (unsure) slot = atomicAdd(&s_n, 1u);
   __syncthreads();
   if (tid == 0 && s_n) s_base = atomicAdd(P.fixup_count, s_n);
   __syncthreads();
   if (unsure) P.fixup_list[s_base + slot] = (uint32_t)i | ((uint32_t)j << 16);
}
/* pass 2: the listed pixels again, in double */
__global__ void __launch_bounds__(128) f3b_final_fixup_kernel(const F3FinalParams P) {
   const unsigned n = *P.fixup_count;
   for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
      const uint32_t e = P.fixup_list[k];
      const int i = (int)(e & 0xffffu), j = (int)(e >> 16);
      f3b_final_pixel<double, double>(P, i, j, f3b_final_origin(P, i, j), (unsigned)P.width, P.filter);
   }
}

int f3b_launch_final(const F3FinalParams &p, cudaStream_t stream) {
   int launches = 0;
   if (p.out_row_end <= p.out_row_begin) return 1;
   if (p.earlyclip) {
      /* the accumulator rows under the filter footprints of the output rows */
      const size_t r0 = (size_t)p.de_offset + (size_t)p.out_row_begin * p.oversample;
      const size_t r1 = (size_t)p.de_offset + (size_t)(p.out_row_end - 1) * p.oversample + p.filter_width;
      const size_t i0 = r0 * p.width, n = r1 * p.width;
      int blocks = (int)((n - i0 + 255) / 256);
      if (blocks > 148 * 16) blocks = 148 * 16;
      f3b_earlyclip_kernel<<<blocks, 256, 0, stream>>>(p, i0, n);
      if (cudaGetLastError() != cudaSuccess) return 0;
      launches++;
   }
   dim3 block(32, 8);
   dim3 grid((p.image_width + 31) / 32, (p.out_row_end - p.out_row_begin + 7) / 8);
   int tile_w = 31 * p.oversample + p.filter_width, tile_h = 7 * p.oversample + p.filter_width;
   size_t smem = 0;
   if (p.fixup_list) {
      const size_t fbytes = ((size_t)p.filter_width * p.filter_width * 4 + 15) & ~(size_t)15;
      smem = fbytes + (size_t)tile_w * tile_h * sizeof(float4);
      if (smem > 96 * 1024) { /* very wide filters: taps straight from the accumulator */
         tile_w = tile_h = 0;
         smem = fbytes;
      }
      if (smem > 200 * 1024) return 0;
      if (cudaFuncSetAttribute(f3b_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
         return 0;
   }
   f3b_final_kernel<<<grid, block, smem, stream>>>(p, tile_w, tile_h);
   if (cudaGetLastError() != cudaSuccess) return 0;
   launches++;
   if (p.fixup_list) {
      f3b_final_fixup_kernel<<<148 * 4, 128, 0, stream>>>(p);
      if (cudaGetLastError() != cudaSuccess) return 0;
      launches++;
   }
   return launches;
}

/* ---------------------------------------------------------------------------------- */
/* histogram import/export in the reference's bits=33 bucket layout (uint32[5])        */
/* ---------------------------------------------------------------------------------- */
__global__ void f3b_hist_from_u32_kernel(const uint32_t *src, void *exact_color, void *exact_count,
                                         int f64, int width, int height, int pitch) {
   const size_t n = (size_t)width * height;
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < n; i += stride) {
      const uint32_t *b = src + i * 5;
      const size_t y = i / (size_t)width, o = (i - y * (size_t)width) + y * (size_t)pitch;
      if (f64) {
         reinterpret_cast<double4 *>(exact_color)[o] =
             make_double4((double)b[0], (double)b[1], (double)b[2], (double)b[3]);
         if (exact_count) reinterpret_cast<double *>(exact_count)[o] = (double)b[4];
      } else {
         reinterpret_cast<uint4 *>(exact_color)[o] = make_uint4(b[0], b[1], b[2], b[3]);
         if (exact_count) reinterpret_cast<uint32_t *>(exact_count)[o] = b[4];
      }
   }
}
__device__ __forceinline__ uint32_t f3b_sat_u32(double v) {
   return v <= 0.0 ? 0u : (v >= 4294967295.0 ? 0xffffffffu : (uint32_t)v);
}
__global__ void f3b_hist_to_u32_kernel(const void *exact_color, const void *exact_count, int f64,
                                       uint32_t *dst, int width, int height, int pitch) {
   const size_t n = (size_t)width * height;
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < n; i += stride) {
      const size_t y = i / (size_t)width, o = (i - y * (size_t)width) + y * (size_t)pitch;
      uint32_t *b = dst + i * 5;
      if (f64) {
         const double4 c = reinterpret_cast<const double4 *>(exact_color)[o];
         b[0] = f3b_sat_u32(c.x);
         b[1] = f3b_sat_u32(c.y);
         b[2] = f3b_sat_u32(c.z);
         b[3] = f3b_sat_u32(c.w);
         b[4] = exact_count ? f3b_sat_u32(reinterpret_cast<const double *>(exact_count)[o]) : b[3];
      } else {
         const uint4 c = reinterpret_cast<const uint4 *>(exact_color)[o];
         b[0] = c.x;
         b[1] = c.y;
         b[2] = c.z;
         b[3] = c.w;
         b[4] = exact_count ? reinterpret_cast<const uint32_t *>(exact_count)[o] : c.w;
      }
   }
}
int f3b_launch_hist_from_u32(const uint32_t *src5, void *exact_color, void *exact_count, int f64,
                             int width, int height, int pitch, cudaStream_t stream) {
   f3b_hist_from_u32_kernel<<<148 * 8, 256, 0, stream>>>(src5, exact_color, exact_count, f64, width,
                                                         height, pitch);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}
int f3b_launch_hist_to_u32(const void *exact_color, const void *exact_count, int f64, uint32_t *dst5,
                           int width, int height, int pitch, cudaStream_t stream) {
   f3b_hist_to_u32_kernel<<<148 * 8, 256, 0, stream>>>(exact_color, exact_count, f64, dst5, width,
                                                       height, pitch);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}

/* ---------------------------------------------------------------------------------- */
/* staging copies -> exact buckets (f3b_iterate.cu: the sweeper warp)                  */
/* ---------------------------------------------------------------------------------- */
/* exact += every staging copy, which is left zeroed for the next launch (so the staging
 * planes are never cleared separately). Integer buckets saturate at UINT_MAX like
 * bump_no_overflow (flam3.c:3838-3854); a staged value found at >= 2^24 is counted as a loss
 * of exactness. Also records the fullest bucket (aggregation decision of the next frame). */
__device__ __forceinline__ uint32_t f3b_merge_u32(uint32_t base, long long staged) {
   const long long s = (long long)base + staged;
   return s <= 0 ? 0u : (s >= 0xffffffffll ? 0xffffffffu : (uint32_t)s);
}
__global__ void __launch_bounds__(256) f3b_hist_merge_kernel(const F3MergeParams M) {
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   unsigned int mx = 0;
   unsigned int lost = 0;
   const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
   for (; i < M.elems; i += stride) {
      if (M.f64) {
         double ax = 0.0, ay = 0.0, az = 0.0, aw = 0.0, ac = 0.0;
         for (uint32_t r = 0; r < M.lay.replicas; r++) {
            const size_t e = f3b_stage_index(M.lay, (uint32_t)i) + r * M.lay.replica_stride;
            const float4 b = M.color[e];
            ax += (double)b.x;
            ay += (double)b.y;
            az += (double)b.z;
            aw += (double)b.w;
            M.color[e] = zero4;
            if (M.count) {
               ac += (double)M.count[e];
               M.count[e] = 0.0f;
            }
         }
         double4 *ec = reinterpret_cast<double4 *>(M.exact_color) + i;
         double4 v = *ec;
         v.x += ax;
         v.y += ay;
         v.z += az;
         v.w += aw;
         *ec = v;
         double cnt = v.w;
         if (M.count) {
            double *en = reinterpret_cast<double *>(M.exact_count) + i;
            cnt = *en + ac;
            *en = cnt;
         }
         mx = max(mx, (unsigned int)fmin(cnt, 4294967295.0));
      } else {
         long long ax = 0, ay = 0, az = 0, aw = 0, ac = 0;
         for (uint32_t r = 0; r < M.lay.replicas; r++) {
            const size_t e = f3b_stage_index(M.lay, (uint32_t)i) + r * M.lay.replica_stride;
            const float4 b = M.color[e];
            const float m = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
            if (m >= F3B_EXACT_LIMIT) lost++;
            if (m != 0.0f) {
               ax += (long long)b.x;
               ay += (long long)b.y;
               az += (long long)b.z;
               aw += (long long)b.w;
               M.color[e] = zero4;
            }
            if (M.count) {
               const float c = M.count[e];
               if (fabsf(c) >= F3B_EXACT_LIMIT) lost++;
               if (c != 0.0f) {
                  ac += (long long)c;
                  M.count[e] = 0.0f;
               }
            }
         }
         uint4 *ec = reinterpret_cast<uint4 *>(M.exact_color) + i;
         uint4 v = *ec;
         if (ax | ay | az | aw) {
            v.x = f3b_merge_u32(v.x, ax);
            v.y = f3b_merge_u32(v.y, ay);
            v.z = f3b_merge_u32(v.z, az);
            v.w = f3b_merge_u32(v.w, aw);
            *ec = v;
         }
         unsigned int cnt = v.w;
         if (M.count) {
            uint32_t *en = reinterpret_cast<uint32_t *>(M.exact_count) + i;
            cnt = *en;
            if (ac) {
               cnt = f3b_merge_u32(cnt, ac);
               *en = cnt;
            }
         }
         mx = max(mx, cnt);
      }
   }
   for (int o = 16; o > 0; o >>= 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      lost += __shfl_xor_sync(0xffffffffu, lost, o);
   }
   if ((threadIdx.x & 31) == 0) {
      if (M.hot_max && mx > 0) atomicMax(M.hot_max, mx);
      if (lost) {
         atomicAdd(M.exact_flags, (unsigned long long)lost);
         atomicAdd(M.exact_flags + 6, (unsigned long long)lost);
      }
   }
}
/* uint32 sums that wrapped while the sweepers were adding to them are pinned to UINT_MAX */
__global__ void f3b_hist_saturate_kernel(const F3MergeParams M) {
   unsigned long long n = M.exact_flags[1];
   if (n > F3B_SAT_LIST_CAP) n = F3B_SAT_LIST_CAP;
   for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
        i += (unsigned long long)gridDim.x * blockDim.x) {
      const uint32_t e = M.sat_list[i];
      const size_t bucket = e >> 3;
      const int ch = (int)(e & 7u);
      if (ch < 4) reinterpret_cast<uint32_t *>(M.exact_color)[bucket * 4 + ch] = 0xffffffffu;
      else if (M.exact_count) reinterpret_cast<uint32_t *>(M.exact_count)[bucket] = 0xffffffffu;
   }
}
int f3b_launch_hist_merge(const F3MergeParams &m, cudaStream_t stream) {
   int blocks = (int)((m.elems + 255) / 256);
   if (blocks > 148 * 8) blocks = 148 * 8;
   f3b_hist_merge_kernel<<<blocks, 256, 0, stream>>>(m);
   if (cudaGetLastError() != cudaSuccess) return -1;
   if (m.f64 || !m.sat_check) return 1;
   f3b_hist_saturate_kernel<<<1, 256, 0, stream>>>(m);
   return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

/* ---------------------------------------------------------------------------------- */
/* strip reduction over peer memory (several GPUs on one frame)                        */
/* ---------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) f3b_hist_reduce_kernel(const F3ReduceParams R) {
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < R.n; i += stride) {
      const size_t e = R.first + i;
      if (R.f64) {
         double4 a = make_double4(0.0, 0.0, 0.0, 0.0);
         double c = 0.0;
#pragma unroll 4
         for (int p = 0; p < R.npeers; p++) {
            const double4 b = reinterpret_cast<const double4 *>(R.color[p])[e];
            a.x += b.x;
            a.y += b.y;
            a.z += b.z;
            a.w += b.w;
            if (R.out_count) c += reinterpret_cast<const double *>(R.count[p])[e];
         }
         reinterpret_cast<double4 *>(R.out_color)[i] = a;
         if (R.out_count) reinterpret_cast<double *>(R.out_count)[i] = c;
      } else {
         unsigned long long ax = 0, ay = 0, az = 0, aw = 0, ac = 0;
#pragma unroll 4
         for (int p = 0; p < R.npeers; p++) {
            const uint4 b = reinterpret_cast<const uint4 *>(R.color[p])[e];
            ax += b.x;
            ay += b.y;
            az += b.z;
            aw += b.w;
            if (R.out_count) ac += reinterpret_cast<const uint32_t *>(R.count[p])[e];
         }
         const unsigned long long top = 0xffffffffull;
         reinterpret_cast<uint4 *>(R.out_color)[i] =
             make_uint4((uint32_t)min(ax, top), (uint32_t)min(ay, top), (uint32_t)min(az, top),
                        (uint32_t)min(aw, top));
         if (R.out_count) reinterpret_cast<uint32_t *>(R.out_count)[i] = (uint32_t)min(ac, top);
      }
   }
}
int f3b_launch_hist_reduce(const F3ReduceParams &r, cudaStream_t stream) {
   if (r.n == 0) return 1;
   int blocks = (int)((r.n + 255) / 256);
   if (blocks > 148 * 8) blocks = 148 * 8;
   f3b_hist_reduce_kernel<<<blocks, 256, 0, stream>>>(r);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}
