/*
 * f3b_filters.c — spatial, temporal and density-estimation filter construction.
 * Behavioural specification: filters.c (kernel functions :47-202, spatial :217-270,
 * DE bank :272-407, temporal :409-489). Runs once per frame / per batch on the host; the
 * tables are then uploaded by f3b_api.cu.
 */
#include "f3b_filters.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define DE_KEEP 100 /* DE_THRESH / keep_thresh, filters.c:281 */

/* kernel support radii in the order of the numeric dispatch (filters.c:29-45) */
static const double k_support[F3B_NUM_SPATIAL_KERNELS] = {
    1.5, 1.0, 0.5, 1.0, 1.5, 2.0, 2.0, 1.0, 2.0, 1.0, 1.0, 3.0, 2.0, 1.5};

static double sinc(double x) {
   x *= M_PI;
   return x != 0.0 ? sin(x) / x : 1.0;
}

static double k_gaussian(double x) { return exp(-2.0 * x * x) * sqrt(2.0 / M_PI); }

static double k_hermite(double t) {
   t = fabs(t);
   return t < 1.0 ? (2.0 * t - 3.0) * t * t + 1.0 : 0.0;
}

static double k_box(double t) { return (t > -0.5 && t <= 0.5) ? 1.0 : 0.0; }

static double k_triangle(double t) {
   t = fabs(t);
   return t < 1.0 ? 1.0 - t : 0.0;
}

static double k_bell(double t) {
   t = fabs(t);
   if (t < 0.5) return 0.75 - t * t;
   if (t < 1.5) {
      t -= 1.5;
      return 0.5 * t * t;
   }
   return 0.0;
}

static double k_bspline(double t) {
   t = fabs(t);
   if (t < 1.0) {
      const double tt = t * t;
      return 0.5 * tt * t - tt + 2.0 / 3.0;
   }
   if (t < 2.0) {
      t = 2.0 - t;
      return t * t * t / 6.0;
   }
   return 0.0;
}

static double k_mitchell(double t) {
   const double B = 1.0 / 3.0, C = 1.0 / 3.0; /* flam3_mitchell_b/c, filters.h */
   const double tt = t * t;
   t = fabs(t);
   if (t < 1.0)
      return ((12.0 - 9.0 * B - 6.0 * C) * (t * tt) + (-18.0 + 12.0 * B + 6.0 * C) * tt +
              (6.0 - 2.0 * B)) / 6.0;
   if (t < 2.0)
      return ((-B - 6.0 * C) * (t * tt) + (6.0 * B + 30.0 * C) * tt +
              (-12.0 * B - 48.0 * C) * t + (8.0 * B + 24.0 * C)) / 6.0;
   return 0.0;
}

static double k_catrom(double x) {
   if (x < -2.0) return 0.0;
   if (x < -1.0) return 0.5 * (4.0 + x * (8.0 + x * (5.0 + x)));
   if (x < 0.0) return 0.5 * (2.0 + x * x * (-5.0 - 3.0 * x));
   if (x < 1.0) return 0.5 * (2.0 + x * x * (-5.0 + 3.0 * x));
   if (x < 2.0) return 0.5 * (4.0 + x * (-8.0 + x * (5.0 - x)));
   return 0.0;
}

static double k_lanczos(double t, double n) {
   t = fabs(t);
   return t < n ? sinc(t) * sinc(t / n) : 0.0;
}

static double k_quadratic(double x) {
   if (x < -1.5) return 0.0;
   if (x < -0.5) return 0.5 * (x + 1.5) * (x + 1.5);
   if (x < 0.5) return 0.75 - x * x;
   if (x < 1.5) return 0.5 * (x - 1.5) * (x - 1.5);
   return 0.0;
}

double f3b_spatial_support(int select) {
   if (select < 0 || select >= F3B_NUM_SPATIAL_KERNELS) select = 0;
   return k_support[select];
}

double f3b_spatial_kernel(int select, double x) {
   switch (select) {
   case 0: return k_gaussian(x);
   case 1: return k_hermite(x);
   case 2: return k_box(x);
   case 3: return k_triangle(x);
   case 4: return k_bell(x);
   case 5: return k_bspline(x);
   case 6: return k_mitchell(x);
   case 7: return sinc(x) * (0.42 + 0.5 * cos(M_PI * x) + 0.08 * cos(2.0 * M_PI * x));
   case 8: return k_catrom(x);
   case 9: return sinc(x) * (0.5 + 0.5 * cos(M_PI * x));
   case 10: return sinc(x) * (0.54 + 0.46 * cos(M_PI * x));
   /* the reference multiplies the windowed sinc by the window once more (filters.c:196-199) */
   case 11: return k_lanczos(x, 3.0) * sinc(x / 3.0);
   case 12: return k_lanczos(x, 2.0) * sinc(x / 2.0);
   default: return k_quadratic(x);
   }
}

int f3b_spatial_filter_create(int select, int oversample, double radius, double pixel_aspect,
                              int field, double **filter) {
   if (select < 0 || select >= F3B_NUM_SPATIAL_KERNELS) select = F3B_NUM_SPATIAL_KERNELS - 1;
   const double supp = k_support[select];
   const double fw = 2.0 * supp * oversample * radius / pixel_aspect;
   int width = (int)fw + 1;
   if ((width ^ oversample) & 1) width++; /* same parity as the oversample factor */
   const double adjust = fw > 0.0 ? supp * width / fw : 1.0;

   double *k = (double *)calloc((size_t)width * width, sizeof(double));
   if (!k) return -1;
   double total = 0.0;
   for (int i = 0; i < width; i++) {
      for (int j = 0; j < width; j++) {
         const double ii = ((2.0 * i + 1.0) / (double)width - 1.0) * adjust;
         double jj = ((2.0 * j + 1.0) / (double)width - 1.0) * adjust;
         if (field) jj *= 2.0; /* scanline fields */
         jj /= pixel_aspect;
         k[i + j * width] = f3b_spatial_kernel(select, ii) * f3b_spatial_kernel(select, jj);
      }
   }
   for (int i = 0; i < width * width; i++) total += k[i];
   if (total == 0.0) {
      free(k);
      return -1;
   }
   total = 1.0 / total;
   for (int i = 0; i < width * width; i++) k[i] *= total;
   *filter = k;
   return width;
}

double f3b_temporal_filter_create(int numsteps, int type, double filter_exp, double filter_width,
                                  double **weights, double **deltas) {
   double *d = (double *)malloc(sizeof(double) * (size_t)numsteps);
   double *w = (double *)malloc(sizeof(double) * (size_t)numsteps);
   *deltas = d;
   *weights = w;
   if (numsteps == 1) {
      d[0] = 0.0;
      w[0] = 1.0;
      return 1.0;
   }
   for (int i = 0; i < numsteps; i++)
      d[i] = ((double)i / (double)(numsteps - 1) - 0.5) * filter_width;

   double maxw = 0.0;
   if (type == 2) { /* exp */
      for (int i = 0; i < numsteps; i++) {
         const double slpx = filter_exp >= 0 ? ((double)i + 1.0) / numsteps
                                             : (double)(numsteps - i) / numsteps;
         w[i] = pow(slpx, fabs(filter_exp));
         if (w[i] > maxw) maxw = w[i];
      }
   } else if (type == 1) { /* gaussian */
      const double half = numsteps / 2.0;
      for (int i = 0; i < numsteps; i++) {
         w[i] = k_gaussian(k_support[0] * fabs(i - half) / half);
         if (w[i] > maxw) maxw = w[i];
      }
   } else { /* box */
      for (int i = 0; i < numsteps; i++) w[i] = 1.0;
      maxw = 1.0;
   }
   double sum = 0.0;
   for (int i = 0; i < numsteps; i++) {
      w[i] /= maxw;
      sum += w[i];
   }
   return sum / numsteps;
}

void f3b_de_filters_free(f3b_de_filters *de) {
   if (!de) return;
   free(de->widths);
   free(de->coefs);
   de->widths = de->coefs = NULL;
}

int f3b_de_filters_create(double max_rad, double min_rad, double curve, int oversample,
                          f3b_de_filters *de) {
   de->kernel_size = -1;
   de->widths = de->coefs = NULL;
   de->max_filter_index = 0;
   de->n_kernels = 0;
   if (curve <= 0.0) {
      fprintf(stderr, "estimator curve must be > 0\n");
      return 1;
   }
   if (max_rad < min_rad) {
      fprintf(stderr, "estimator must be larger than estimator_minimum.\n");
      fprintf(stderr, "(%f > %f) ? \n", max_rad, min_rad);
      return 1;
   }
   /* radii in buckets; +1 is the assumed distance to the first neighbour */
   const double rmax = max_rad * oversample + 1.0;
   const double rmin = min_rad * oversample + 1.0;

   const double nd = pow(rmax / rmin, 1.0 / curve);
   if (nd > 1e7) {
      fprintf(stderr, "too many filters required in this configuration (%g)\n", nd);
      return 1;
   }
   const int nfilters = (int)ceil(nd);
   int nrows;
   if (nfilters > DE_KEEP) { /* kernels past 100 hits are spaced on the decay curve */
      nrows = (int)ceil(DE_KEEP + pow(nfilters - DE_KEEP, curve)) + 1;
      de->max_filtered_counts = (int)pow((double)(nrows - DE_KEEP), 1.0 / curve) + DE_KEEP;
   } else {
      nrows = nfilters;
      de->max_filtered_counts = nrows;
   }
   const int row = (int)(2 * ceil(rmax) - 1);
   const int half = (row - 1) / 2;
   de->kernel_size = (half + 1) * (2 + half) / 2;
   de->n_kernels = nrows;
   de->coefs = (double *)calloc((size_t)nrows * de->kernel_size, sizeof(double));
   de->widths = (double *)calloc((size_t)nrows, sizeof(double));
   if (!de->coefs || !de->widths) {
      f3b_de_filters_free(de);
      de->kernel_size = -1;
      return 1;
   }

   for (int f = 0; f < nrows; f++) {
      double h;
      if (f < DE_KEEP)
         h = rmax / pow(f + 1, curve);
      else {
         const double adj = pow(f - DE_KEEP, 1.0 / curve) + DE_KEEP;
         h = rmax / pow(adj + 1, curve);
      }
      if (h <= rmin) { /* reached the minimum radius: this is the last kernel */
         h = rmin;
         de->max_filter_index = f;
      }
      de->widths[f] = h;

      /* normalise over the whole (2*half+1)^2 square, store one octant */
      double sum = 0.0;
      for (int j = -half; j <= half; j++)
         for (int k = -half; k <= half; k++) {
            const double d = sqrt((double)(j * j + k * k)) / h;
            if (d <= 1.0) sum += k_gaussian(k_support[0] * d);
         }
      double *c = de->coefs + (size_t)f * de->kernel_size;
      for (int j = 0; j <= half; j++)
         for (int k = 0; k <= j; k++) {
            const double d = sqrt((double)(j * j + k * k)) / h;
            *c++ = d > 1.0 ? 0.0 : k_gaussian(k_support[0] * d) / sum;
         }
      if (de->max_filter_index > 0) break;
   }
   if (de->max_filter_index == 0) de->max_filter_index = nrows - 1;
   return 0;
}
