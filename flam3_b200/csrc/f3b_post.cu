/*
 * f3b_post.cu — the post-pass for sm_100a: log-density scaling + variable-width density
 * estimation, early-clip, spatial filter + tone map + quantisation.
 *
 * Replaces the reference's de_thread (rect.c:57-251), the no-DE accumulate loop
 * (rect.c:947-974), the early-clip sweep (rect.c:1086-1135) and the final filter loop
 * (rect.c:1137-1247) with palettes.c:199-349 (calc_alpha / calc_newrgb / hsv).
 *
 * Density estimation is reformulated from the reference's per-source SCATTER (racy across
 * thread swaths, rect.c:148-168) into a race-free per-target GATHER: a CTA stages a tile of
 * buckets plus a halo of the maximum kernel radius in shared memory, converts every staged
 * bucket once to (log-scaled colour, kernel index), then each thread sums the kernels of
 * the neighbours that reach it. The gather window is clamped to the largest radius
 * actually present in the staged region, so dense regions (radius 0-2) cost a few taps
 * while sparse ones pay for their wide kernels, as in the reference.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "f3b_internal.h"

#define DE_TX 32
#define DE_TY 16
#define DE_THREADS (DE_TX * DE_TY)
#define DE_EMPTY 0xffu

/* ---------------------------------------------------------------------------------- */
/* log scale without density estimation (rect.c:947-974)                               */
/* ---------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) f3b_logscale_kernel(const F3DensityParams P, size_t i0,
                                                           size_t n) {
   size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < n; i += stride) {
      const float4 c = P.hist_color[i];
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
/* density estimation, gather form                                                     */
/* ---------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(DE_THREADS) f3b_density_kernel(const F3DensityParams P) {
   extern __shared__ __align__(16) unsigned char de_smem[];
   const int R = P.max_radius;
   const int os = P.oversample;
   const int ss = os / 2; /* floor(oversample/2), rect.c:61 */
   const int SW = DE_TX + 2 * R, SH = DE_TY + 2 * R; /* staged bucket region */
   const int CW = SW + 2 * ss, CH = SH + 2 * ss;     /* staged count region  */

   float4 *s_col = reinterpret_cast<float4 *>(de_smem);
   float *s_cnt = reinterpret_cast<float *>(s_col + (size_t)SW * SH);
   uint16_t *s_kidx = reinterpret_cast<uint16_t *>(s_cnt + (size_t)CW * CH);
   uint8_t *s_rad = reinterpret_cast<uint8_t *>(s_kidx + (size_t)SW * SH);
   __shared__ int s_maxrad;

   const int tid = threadIdx.y * DE_TX + threadIdx.x;
   const int x0 = blockIdx.x * DE_TX - R; /* bucket coords of staged (0,0) */
   const int y0 = P.row_begin + blockIdx.y * DE_TY - R;
   const bool have_count = (P.hist_count != nullptr);

   if (tid == 0) s_maxrad = 0;

   /* stage counts (only needed separately when oversample > 1 or a count plane exists) */
   for (int i = tid; i < CW * CH; i += DE_THREADS) {
      const int cx = x0 - ss + (i % CW), cy = y0 - ss + (i / CW);
      float c = 0.0f;
      if (cx >= 0 && cx < P.width && cy >= 0 && cy < P.height) {
         const size_t gi = (size_t)cx + (size_t)P.width * cy;
         c = have_count ? P.hist_count[gi] : P.hist_color[gi].w;
      }
      s_cnt[i] = c;
   }
   __syncthreads();

   /* prep: every staged bucket -> (scaled colour, kernel index, radius) */
   const double scfact = (os & 1) ? 1.0 : ((double)os / (os + 1.0)) * ((double)os / (os + 1.0));
   int my_maxrad = 0;
   for (int i = tid; i < SW * SH; i += DE_THREADS) {
      const int lx = i % SW, ly = i / SW;
      const int gx = x0 + lx, gy = y0 + ly;
      float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
      unsigned rad = DE_EMPTY;
      unsigned kidx = 0;
      /* source range of the reference loops (rect.c:79-80), minus its out-of-range row */
      if (gx >= os - 1 && gx <= P.width - os && gy >= os - 1 && gy <= P.height - os) {
         const size_t gi = (size_t)gx + (size_t)P.width * gy;
         const float4 c = P.hist_color[gi];
         const float cnt = s_cnt[(lx + ss) + (ly + ss) * CW];
         if (cnt != 0.0f && c.w != 0.0f) {
            double f = 0.0;
            for (int jj = -ss; jj <= ss; jj++)
               for (int ii = -ss; ii <= ss; ii++)
                  f += (double)s_cnt[(lx + ss + ii) + (ly + ss + jj) * CW];
            f *= (1.0 / 255.0);
            f *= scfact;
            int k;
            if (f > (double)P.max_filtered_counts) k = P.max_filter_index;
            else if (f <= 100.0) k = (int)ceil(f) - 1;
            else k = 100 + (int)floor(pow(f - 100.0, (double)P.curve));
            if (k > P.max_filter_index) k = P.max_filter_index;
            if (k < 0) k = 0;
            kidx = (unsigned)k;
            rad = P.de_radius[k];
            const float ls = (P.k1 * log1pf(c.w * P.k2)) / c.w;
            col = make_float4(c.x * ls, c.y * ls, c.z * ls, c.w * ls);
            my_maxrad = max(my_maxrad, (int)rad);
         }
      }
      s_col[i] = col;
      s_kidx[i] = (uint16_t)kidx;
      s_rad[i] = (uint8_t)rad;
   }
   if (my_maxrad) atomicMax(&s_maxrad, my_maxrad);
   __syncthreads();

   const int gx = blockIdx.x * DE_TX + threadIdx.x;
   const int gy = P.row_begin + blockIdx.y * DE_TY + threadIdx.y;
   if (gx >= P.width || gy >= P.row_end) return;

   const int Rb = s_maxrad;
   const int cx = threadIdx.x + R, cy = threadIdx.y + R;
   float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
   for (int dy = -Rb; dy <= Rb; dy++) {
      const int ady = abs(dy);
      const int rowbase = (cy + dy) * SW + cx;
      for (int dx = -Rb; dx <= Rb; dx++) {
         const int n = rowbase + dx;
         const unsigned rn = s_rad[n];
         const int adx = abs(dx);
         const int hi = max(adx, ady), lo = min(adx, ady);
         if (rn == DE_EMPTY || hi > (int)rn) continue;
         /* octant-packed coefficient: index jj*(jj+1)/2 + ii with jj >= ii (filters.c:362) */
         const float coef =
             __ldg(P.de_coefs + (size_t)s_kidx[n] * P.kernel_size + (hi * (hi + 1)) / 2 + lo);
         const float4 c = s_col[n];
         acc.x += coef * c.x;
         acc.y += coef * c.y;
         acc.z += coef * c.z;
         acc.w += coef * c.w;
      }
   }
   const size_t gi = (size_t)gx + (size_t)P.width * gy;
   float4 a = P.accum[gi];
   a.x += acc.x;
   a.y += acc.y;
   a.z += acc.z;
   a.w += acc.w;
   P.accum[gi] = a;
}

int f3b_launch_density(const F3DensityParams &p, cudaStream_t stream) {
   if (p.row_end <= p.row_begin) return 1;
   if (p.n_kernels <= 0) {
      const size_t i0 = (size_t)p.width * p.row_begin, n = (size_t)p.width * p.row_end;
      int blocks = (int)((n - i0 + 255) / 256);
      if (blocks > 148 * 16) blocks = 148 * 16;
      f3b_logscale_kernel<<<blocks, 256, 0, stream>>>(p, i0, n);
      return cudaGetLastError() == cudaSuccess ? 1 : 0;
   }
   const int R = p.max_radius, ss = p.oversample / 2;
   const size_t SW = DE_TX + 2 * R, SH = DE_TY + 2 * R;
   const size_t CW = SW + 2 * ss, CH = SH + 2 * ss;
   size_t smem = SW * SH * sizeof(float4) + CW * CH * sizeof(float) + SW * SH * 2 + SW * SH;
   smem = (smem + 15) & ~(size_t)15;
   if (smem > 220 * 1024) return 0;
   if (cudaFuncSetAttribute(f3b_density_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)smem) != cudaSuccess)
      return 0;
   dim3 grid((p.width + DE_TX - 1) / DE_TX, (p.row_end - p.row_begin + DE_TY - 1) / DE_TY);
   dim3 block(DE_TX, DE_TY);
   f3b_density_kernel<<<grid, block, smem, stream>>>(p);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}

/* ---------------------------------------------------------------------------------- */
/* tone map helpers (palettes.c:199-349), double precision                             */
/* ---------------------------------------------------------------------------------- */
__device__ __forceinline__ double f3b_calc_alpha(double dens, double g, double linrange) {
   if (!(dens > 0.0)) return 0.0;
   if (dens < linrange) {
      const double funcval = pow(linrange, g);
      const double frac = dens / linrange;
      return (1.0 - frac) * dens * (funcval / linrange) + frac * pow(dens, g);
   }
   return pow(dens, g);
}

__device__ void f3b_rgb2hsv(const double *rgb, double *hsv) {
   const double r = rgb[0], g = rgb[1], b = rgb[2];
   const double mx = fmax(r, fmax(g, b)), mn = fmin(r, fmin(g, b));
   const double del = mx - mn;
   const double s = (mx != 0.0) ? del / mx : 0.0;
   double h = 0.0;
   if (s != 0.0) {
      const double rc = (mx - r) / del, gc = (mx - g) / del, bc = (mx - b) / del;
      if (r == mx) h = bc - gc;
      else if (g == mx) h = 2.0 + rc - bc;
      else if (b == mx) h = 4.0 + gc - rc;
      if (h < 0.0) h += 6.0;
   }
   hsv[0] = h;
   hsv[1] = s;
   hsv[2] = mx;
}

__device__ void f3b_hsv2rgb(const double *hsv, double *rgb) {
   double h = hsv[0];
   const double s = hsv[1], v = hsv[2];
   while (h >= 6.0) h -= 6.0;
   while (h < 0.0) h += 6.0;
   const int j = (int)floor(h);
   const double f = h - j;
   const double p = v * (1.0 - s), q = v * (1.0 - s * f), t = v * (1.0 - s * (1.0 - f));
   double r, g, b;
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

__device__ void f3b_calc_newrgb(const double *cbuf, double ls, double highpow, double *newrgb) {
   if (ls == 0.0 || (cbuf[0] == 0.0 && cbuf[1] == 0.0 && cbuf[2] == 0.0)) {
      newrgb[0] = newrgb[1] = newrgb[2] = 0.0;
      return;
   }
   /* most saturated channel */
   double maxa = -1.0, maxc = 0.0;
   for (int i = 0; i < 3; i++) {
      const double a = ls * (cbuf[i] / 255.0);
      if (a > maxa) {
         maxa = a;
         maxc = cbuf[i] / 255.0;
      }
   }
   const double newls = 255.0 / maxc;
   if (maxa > 255.0 && highpow >= 0.0) {
      /* clip to full scale and desaturate instead of shifting hue */
      const double lsratio = pow(newls / ls, highpow);
      double hsv[3];
      for (int i = 0; i < 3; i++) newrgb[i] = newls * (cbuf[i] / 255.0) / 255.0;
      f3b_rgb2hsv(newrgb, hsv);
      hsv[1] *= lsratio;
      f3b_hsv2rgb(hsv, newrgb);
      for (int i = 0; i < 3; i++) newrgb[i] *= 255.0;
   } else {
      double adj = -highpow;
      if (adj > 1.0) adj = 1.0;
      if (maxa <= 255.0) adj = 1.0;
      for (int i = 0; i < 3; i++)
         newrgb[i] = ((1.0 - adj) * newls + adj * ls) * (cbuf[i] / 255.0);
   }
}

/* gamma / vibrancy / highlight / background for one 4-vector t (rect.c:1100-1130 and
 * rect.c:1172-1210). On return t[0..2] are clamped to [0,255] and t[3] is alpha. */
__device__ __forceinline__ void f3b_tonemap(double *t, const F3FinalParams &P) {
   double alpha, ls;
   if (t[3] <= 0.0) {
      alpha = 0.0;
      ls = 0.0;
   } else {
      const double tmp = t[3] / 255.0;
      alpha = f3b_calc_alpha(tmp, P.g, P.linrange);
      ls = P.vibrancy * 256.0 * alpha / tmp;
      if (alpha < 0.0) alpha = 0.0;
      if (alpha > 1.0) alpha = 1.0;
   }
   double newrgb[3];
   f3b_calc_newrgb(t, ls, P.highpow, newrgb);
   for (int i = 0; i < 3; i++) {
      double a = newrgb[i];
      a += (1.0 - P.vibrancy) * 256.0 * pow(t[i] / 255.0, P.g);
      if (P.nchan <= 3 || P.transp == 0) a += (1.0 - alpha) * P.background[i];
      else a = (alpha > 0.0) ? a / alpha : 0.0;
      if (a > 255.0) a = 255.0;
      if (a < 0.0) a = 0.0;
      t[i] = a;
   }
   t[3] = alpha;
}

__global__ void __launch_bounds__(256) f3b_earlyclip_kernel(const F3FinalParams P, size_t n) {
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < n; i += stride) {
      const float4 a = P.accum[i];
      double t[4] = {(double)a.x, (double)a.y, (double)a.z, (double)a.w};
      f3b_tonemap(t, P);
      P.accum[i] = make_float4((float)t[0], (float)t[1], (float)t[2], (float)t[3]);
   }
}

__global__ void __launch_bounds__(256) f3b_final_kernel(const F3FinalParams P) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x; /* output column */
   const int j = blockIdx.y * blockDim.y + threadIdx.y; /* output row    */
   if (i >= P.image_width || j >= P.image_height) return;
   const int x = P.de_offset + i * P.oversample;
   const int y = P.de_offset + j * P.oversample;
   const int fw = P.filter_width;

   double t[4] = {0.0, 0.0, 0.0, 0.0};
   /* same tap order as rect.c:1146-1160 (ii outer, jj inner) so the double sums agree */
   for (int ii = 0; ii < fw; ii++) {
      for (int jj = 0; jj < fw; jj++) {
         const double k = __ldg(P.filter + ii + jj * fw);
         const float4 a = __ldg(P.accum + (size_t)(x + ii) + (size_t)(y + jj) * P.width);
         t[0] += k * (double)a.x;
         t[1] += k * (double)a.y;
         t[2] += k * (double)a.z;
         t[3] += k * (double)a.w;
      }
   }
   if (!P.earlyclip) f3b_tonemap(t, P);

   double alpha = t[3];
   if (alpha > 1.0) alpha = 1.0;
   if (alpha < 0.0) alpha = 0.0;
   const size_t pix = (size_t)i + (size_t)j * P.image_width;
   if (P.bytes_per_channel == 2) {
      unsigned short *o = reinterpret_cast<unsigned short *>(P.out) + pix * P.nchan;
      for (int c = 0; c < 3; c++) {
         double a = t[c];
         if (a > 255.0) a = 255.0;
         if (a < 0.0) a = 0.0;
         o[c] = (unsigned short)(a * 256.0);
      }
      if (P.nchan > 3) o[3] = P.transp ? (unsigned short)(alpha * 65535.0) : 65535;
   } else {
      unsigned char *o = reinterpret_cast<unsigned char *>(P.out) + pix * P.nchan;
      for (int c = 0; c < 3; c++) {
         double a = t[c];
         if (a > 255.0) a = 255.0;
         if (a < 0.0) a = 0.0;
         o[c] = (unsigned char)a;
      }
      if (P.nchan > 3) o[3] = P.transp ? (unsigned char)(alpha * 255.0) : 255;
   }
}

int f3b_launch_final(const F3FinalParams &p, cudaStream_t stream) {
   int launches = 0;
   if (p.earlyclip) {
      const size_t n = (size_t)p.width * p.height;
      int blocks = (int)((n + 255) / 256);
      if (blocks > 148 * 16) blocks = 148 * 16;
      f3b_earlyclip_kernel<<<blocks, 256, 0, stream>>>(p, n);
      if (cudaGetLastError() != cudaSuccess) return 0;
      launches++;
   }
   dim3 block(32, 8);
   dim3 grid((p.image_width + 31) / 32, (p.image_height + 7) / 8);
   f3b_final_kernel<<<grid, block, 0, stream>>>(p);
   if (cudaGetLastError() != cudaSuccess) return 0;
   return launches + 1;
}

/* ---------------------------------------------------------------------------------- */
/* histogram import/export in the reference's bits=33 bucket layout (uint32[5])        */
/* ---------------------------------------------------------------------------------- */
__global__ void f3b_hist_from_u32_kernel(const uint32_t *src, float4 *color, float *count,
                                         size_t n) {
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < n; i += stride) {
      const uint32_t *b = src + i * 5;
      color[i] = make_float4((float)b[0], (float)b[1], (float)b[2], (float)b[3]);
      if (count) count[i] = (float)b[4];
   }
}
__global__ void f3b_hist_to_u32_kernel(const float4 *color, const float *count, uint32_t *dst,
                                       size_t n) {
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t stride = (size_t)gridDim.x * blockDim.x;
   for (; i < n; i += stride) {
      const float4 c = color[i];
      uint32_t *b = dst + i * 5;
      b[0] = (uint32_t)fminf(c.x, 4294967040.0f);
      b[1] = (uint32_t)fminf(c.y, 4294967040.0f);
      b[2] = (uint32_t)fminf(c.z, 4294967040.0f);
      b[3] = (uint32_t)fminf(c.w, 4294967040.0f);
      b[4] = count ? (uint32_t)fminf(count[i], 4294967040.0f) : b[3];
   }
}
int f3b_launch_hist_from_u32(const uint32_t *src5, float4 *color, float *count, size_t n,
                             cudaStream_t stream) {
   f3b_hist_from_u32_kernel<<<148 * 8, 256, 0, stream>>>(src5, color, count, n);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}
int f3b_launch_hist_to_u32(const float4 *color, const float *count, uint32_t *dst5, size_t n,
                           cudaStream_t stream) {
   f3b_hist_to_u32_kernel<<<148 * 8, 256, 0, stream>>>(color, count, dst5, n);
   return cudaGetLastError() == cudaSuccess ? 1 : 0;
}
