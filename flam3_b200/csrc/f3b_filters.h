/*
 * f3b_filters.h — host-side construction of the three filter families of the render path
 * (reference: filters.c). Plain C, double precision, no flam3.h dependency: callers pass
 * the handful of genome fields each filter depends on.
 */
#ifndef F3B_FILTERS_H
#define F3B_FILTERS_H

#ifdef __cplusplus
extern "C" {
#endif

#define F3B_NUM_SPATIAL_KERNELS 14

/* Spatial (down-sampling) filter, filters.c:217-270. `select` follows the reference's
 * NUMERIC dispatch (filters.c:172-202): 0 gaussian, 1 hermite, 2 box, 3 triangle, 4 bell,
 * 5 b-spline, 6 mitchell, 7 blackman, 8 catrom, 9 hanning, 10 hamming, 11 lanczos3,
 * 12 lanczos2, 13 quadratic. Returns the filter width (>0) and a malloc'd width*width
 * kernel normalised to sum 1 in *filter, or -1 when the kernel sums to zero. */
int f3b_spatial_filter_create(int select, int oversample, double radius, double pixel_aspect,
                              int field, double **filter);
double f3b_spatial_kernel(int select, double x);
double f3b_spatial_support(int select);

/* Temporal filter, filters.c:409-489. type: 0 box, 1 gaussian, 2 exp. Fills malloc'd
 * weights[numsteps] (max 1) and deltas[numsteps]; returns the mean weight (sumfilt). */
double f3b_temporal_filter_create(int numsteps, int type, double filter_exp, double filter_width,
                                  double **weights, double **deltas);

/* Density-estimation kernel bank, filters.c:272-407. */
typedef struct {
   int max_filtered_counts;
   int max_filter_index;
   int kernel_size; /* coefficients per kernel (octant), -1 on error */
   int n_kernels;   /* rows allocated (de_max_ind) */
   double *widths;  /* [n_kernels] */
   double *coefs;   /* [n_kernels * kernel_size] */
} f3b_de_filters;
int f3b_de_filters_create(double max_rad, double min_rad, double curve, int oversample,
                          f3b_de_filters *de);
void f3b_de_filters_free(f3b_de_filters *de);

#ifdef __cplusplus
}
#endif
#endif
