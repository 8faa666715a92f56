/*
 * f3b-render — a flame renderer made of this repository's code only: front end
 * (f3b_frontend.c: genome XML, palette library, PPM/PNG), the flam3_render drop-in
 * (f3b_render.c, f3b_pack.c, f3b_isaac.c) and libflam3_b200.so. No reference object is linked
 * (SURVEY.md §8(f) row 4). It takes flam3-render's environment (README.txt:55-115, the subset
 * that matters for stills):
 *
 *    in (default stdin), prefix, out, format (ppm | png), qs, ss, bits (33), bpc (8 | 16, 16 needs png),
 *    nthreads, sub_batch_size, earlyclip, pixel_aspect, transparency, isaac_seed, verbose, nstrips,
 *    write_genome (animation frames),
 *    flam3_palettes, F3B_DEVICE / F3B_DEVICES
 *
 *    env format=ppm prefix=out/ tools/_build/f3b-render < some.flam3
 *
 * One image per <flame>, named <prefix>%05d.<format> like flam3-render.c:355. For the same
 * genome file and isaac_seed the bytes equal those of the unmodified flam3-render linked
 * against the drop-in (tests/test_gpu_frontend.py).
 *
 * With begin / end / time / frame / dtime in the environment it is flam3-animate instead
 * (flam3-animate.c:160-290, without fields): the flames are key frames, one motion-blurred
 * image per time step from `begin` (default: the first key frame's time) to `end` (default: the
 * last one's - 1), named <prefix>%05d.<format> by time; the genomes between the key frames
 * come from this front end's own interpolation (f3b_interpolate).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "flam3.h"

#include "../include/flam3_b200_frontend.h"

/* shims/imgshim.c: this repository's PNG writer behind the reference's write_png signature
 * (img.h:29-40) */
typedef struct {
   char *genome, *badvals, *numiters, *rtime;
} flam3_img_comments;
void write_png(FILE *file, void *image, int width, int height, flam3_img_comments *fpc, int bpc);

static const char *env_or(const char *name, const char *dflt) {
   const char *v = getenv(name);
   return v ? v : dflt;
}

/* the image file of a rendered frame */
static int write_image(flam3_frame *f, const void *image, const stat_struct *stats, int width, int height, int channels,
                       const char *format, const char *name, int verbose) {
   /* the genome that goes into the PNG's comment: the flame itself, or for an animation frame the
    * one interpolated at the frame's time (flam3-render.c:367, flam3-animate.c:287-291) */
   char *genome_text = NULL;
   const int write_genome = f->ngenomes > 1 && atoi(env_or("write_genome", "0")); /* flam3-animate.c:264-279 */
   if (strcmp(format, "ppm") || write_genome) {
      flam3_genome centre;
      memset(&centre, 0, sizeof(centre));
      if (f->ngenomes == 1) genome_text = f3b_print_flam3(&f->genomes[0], NULL, NULL);
      else if (!f3b_interpolate(f->genomes, f->ngenomes, f->time, 0.0, &centre)) {
         genome_text = f3b_print_flam3(&centre, NULL, NULL);
         f3b_genome_free_xforms(&centre);
      }
   }
   if (verbose) fprintf(stderr, "writing %s...", name);
   if (write_genome && genome_text) { /* the frame's interpolated genome beside the image */
      char gname[4200];
      snprintf(gname, sizeof(gname), "%s.flam3", name);
      FILE *gf = fopen(gname, "w");
      if (!gf) {
         perror(gname);
         return 1;
      }
      fputs(genome_text, gf);
      fclose(gf);
   }
   if (!strcmp(format, "ppm")) {
      if (f3b_write_ppm(name, image, width, height, channels, f->bytes_per_channel)) return 1;
   } else {
      FILE *fp = fopen(name, "wb");
      if (!fp) {
         perror(name);
         return 1;
      }
      flam3_img_comments fpc;
      memset(&fpc, 0, sizeof(fpc));
      char badvals[64], iters[64], secs[64];
      snprintf(badvals, sizeof(badvals), "%g", stats->badvals / (double)(stats->num_iters ? stats->num_iters : 1));
      snprintf(iters, sizeof(iters), "%g", (double)stats->num_iters);
      snprintf(secs, sizeof(secs), "%d", stats->render_seconds);
      fpc.genome = genome_text ? genome_text : (char *)"";
      fpc.badvals = badvals;
      fpc.numiters = iters;
      fpc.rtime = secs;
      write_png(fp, (void *)image, width, height, &fpc, f->bytes_per_channel);
      fclose(fp);
   }
   if (verbose) fprintf(stderr, "done.\n");
   free(genome_text);
   return 0;
}
/* one flam3_render call and its image file */
static int render_and_write(flam3_frame *f, int width, int height, int channels, int transp, const char *format,
                            const char *name, int verbose) {
   const size_t bytes = (size_t)channels * width * height * f->bytes_per_channel;
   void *image = calloc(bytes, 1);
   stat_struct stats;
   if (!image || flam3_render(f, image, flam3_field_both, channels, transp, &stats)) {
      fprintf(stderr, "error rendering image: aborting.\n");
      return 1;
   }
   const int rc = write_image(f, image, &stats, width, height, channels, format, name, verbose);
   free(image);
   return rc;
}

int main(void) {
   const char *inf = getenv("in"), *prefix = env_or("prefix", ""), *out = getenv("out");
   const char *format = env_or("format", "ppm");
   const double qs = atof(env_or("qs", "1")), ss = atof(env_or("ss", "1"));
   const int bits = atoi(env_or("bits", "33")), bpc = atoi(env_or("bpc", "8"));
   const int transp = atoi(env_or("transparency", "0")), verbose = atoi(env_or("verbose", "0"));
   if (strcmp(format, "ppm") && strcmp(format, "png")) {
      fprintf(stderr, "f3b-render: format must be ppm or png, not %s.\n", format);
      return 1;
   }
   if (bpc != 8 && bpc != 16) {
      fprintf(stderr, "Unexpected bpc specified (%d)\n", bpc);
      return 1;
   }
   if (bpc == 16 && strcmp(format, "png")) {
      fprintf(stderr, "Support for 16 bpc images is only present for the png format.\n");
      return 1;
   }
   FILE *in = inf ? fopen(inf, "rb") : stdin;
   if (!in) {
      perror(inf);
      return 1;
   }
   size_t cap = 1 << 16, n = 0;
   char *text = (char *)malloc(cap);
   for (size_t got; text && (got = fread(text + n, 1, cap - n - 1, in)) > 0;) {
      n += got;
      if (n + 1 >= cap) text = (char *)realloc(text, cap *= 2);
   }
   if (inf) fclose(in);
   if (!text) return 1;
   text[n] = 0;

   int ncps = 0;
   flam3_genome *cps = f3b_parse_flam3(text, flam3_defaults_on, &ncps);
   free(text);
   if (!cps || ncps == 0) {
      fprintf(stderr, "error reading genomes from file\n");
      return 1;
   }
   flam3_frame f;
   memset(&f, 0, sizeof(f));
   flam3_init_frame(&f);
   const int channels = strcmp(format, "png") ? 3 : 4;
   f.verbose = verbose;
   f.bits = bits;
   f.pixel_aspect_ratio = atof(env_or("pixel_aspect", "1"));
   f.progress = 0;
   f.nthreads = atoi(env_or("nthreads", "0"));
   if (f.nthreads <= 0) f.nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN); /* flam3_count_nthreads, flam3.c:1415 */
   if (f.nthreads < 1) f.nthreads = 1;
   f.earlyclip = atoi(env_or("earlyclip", "0"));
   f.sub_batch_size = atoi(env_or("sub_batch_size", "10000"));
   f.bytes_per_channel = bpc == 16 ? 2 : 1;
   if (getenv("begin") || getenv("end") || getenv("time") || getenv("frame") || getenv("dtime")) {
      /* flam3-animate.c:153-290 */
      for (int i = 0; i < ncps; i++) {
         cps[i].sample_density *= qs;
         cps[i].height = (int)(cps[i].height * ss);
         cps[i].width = (int)(cps[i].width * ss);
         cps[i].pixels_per_unit *= ss;
         if (cps[i].height <= 0 || cps[i].width <= 0) {
            fprintf(stderr, "output image has dimension <=0, aborting.\n");
            return 1;
         }
         if (i > 0 && cps[i].time <= cps[i - 1].time) {
            fprintf(stderr, "error: control points must be sorted by time, but %g <= %g, index %d.\n", cps[i].time,
                    cps[i - 1].time, i);
            return 1;
         }
         if (cps[i].width != cps[0].width || cps[i].height != cps[0].height) {
            fprintf(stderr, "warning: flame %d at time %g size mismatch.  (%d,%d) should be (%d,%d).\n", i, cps[i].time,
                    cps[i].width, cps[i].height, cps[0].width, cps[0].height);
            cps[i].width = cps[0].width;
            cps[i].height = cps[0].height;
         }
      }
      const char *one = getenv("time") ? getenv("time") : getenv("frame");
      int first = atoi(env_or("begin", "0")), last = atoi(env_or("end", "0"));
      const int dtime = atoi(env_or("dtime", "1"));
      if (one) first = last = atoi(one);
      else {
         if (!getenv("begin")) first = (int)cps[0].time;
         if (!getenv("end")) {
            last = (int)cps[ncps - 1].time - 1;
            if (last < first) last = first;
         }
      }
      if (dtime < 1) {
         fprintf(stderr, "dtime must be positive, not %d.\n", dtime);
         return 1;
      }
      f.genomes = cps;
      f.ngenomes = ncps;
      for (int t = first; t <= last; t += dtime) {
         f.time = (double)t;
         if (verbose && (last - first) / dtime >= 1) fprintf(stderr, "time = %d/%d/%d\n", t, last, dtime);
         char name[4096];
         if (out) snprintf(name, sizeof(name), "%s", out);
         else snprintf(name, sizeof(name), "%s%05d.%s", prefix, t, format);
         if (render_and_write(&f, cps[0].width, cps[0].height, channels, transp, format, name, verbose)) return 1;
      }
      f3b_free_genomes(cps, ncps);
      return 0;
   }
   for (int i = 0; i < ncps; i++) {
      cps[i].ntemporal_samples = 1; /* stills: flam3-render.c:235 */
      cps[i].sample_density *= qs;
      cps[i].height = (int)(cps[i].height * ss);
      cps[i].width = (int)(cps[i].width * ss);
      cps[i].pixels_per_unit *= ss;
      if (cps[i].height <= 0 || cps[i].width <= 0) {
         fprintf(stderr, "output image has dimension <=0, aborting.\n");
         return 1;
      }
      f.genomes = &cps[i];
      f.ngenomes = 1;
      f.time = 0.0;
      char name[4096];
      if (out) snprintf(name, sizeof(name), "%s", out);
      else snprintf(name, sizeof(name), "%s%05d.%s", prefix, i, format);
      const int nstrips = atoi(env_or("nstrips", "1"));
      if (nstrips <= 1) {
         if (render_and_write(&f, cps[i].width, cps[i].height, channels, transp, format, name, verbose)) return 1;
         continue;
      }
      /* nstrips (flam3-render.c:277-345): the frame as horizontal strips, each one a render of a
       * genome with the strip's height and centre and nstrips times the quality, all from the
       * same state of the frame's random context */
      if (nstrips > cps[i].height) {
         fprintf(stderr, "cannot have more strips than rows but %d>%d.\n", nstrips, cps[i].height);
         return 1;
      }
      const int full_h = cps[i].height, strip_h = (int)ceil(full_h / (double)nstrips);
      const double cy = cps[i].center[1], ppu = cps[i].pixels_per_unit * pow(2.0, cps[i].zoom), q = cps[i].sample_density;
      const size_t row_bytes = (size_t)cps[i].width * channels * f.bytes_per_channel;
      char *image = (char *)calloc(row_bytes, (size_t)full_h);
      if (!image) return 1;
      const double first_centre = cy - ((nstrips - 1) * strip_h) / (2 * ppu);
      const randctx rc0 = f.rc;
      stat_struct stats, total;
      memset(&total, 0, sizeof(total));
      cps[i].sample_density = q * nstrips;
      for (int strip = 0; strip < nstrips; strip++) {
         cps[i].height = strip_h;
         cps[i].center[1] = first_centre + strip_h * (double)strip / ppu;
         if (strip_h * (strip + 1) > full_h) { /* the last strip may be shorter */
            cps[i].height = full_h - strip_h * strip;
            cps[i].center[1] -= (strip_h - cps[i].height) * 0.5 / ppu;
         }
         f.rc = rc0;
         if (verbose) fprintf(stderr, "strip = %d/%d\n", strip + 1, nstrips);
         if (flam3_render(&f, image + (size_t)strip_h * strip * row_bytes, flam3_field_both, channels, transp, &stats)) {
            fprintf(stderr, "error rendering image: aborting.\n");
            return 1;
         }
         total.badvals += stats.badvals;
         total.num_iters += stats.num_iters;
         total.render_seconds += stats.render_seconds;
      }
      cps[i].sample_density = q;
      cps[i].height = full_h;
      cps[i].center[1] = cy;
      if (write_image(&f, image, &total, cps[i].width, full_h, channels, format, name, verbose)) return 1;
      free(image);
   }
   f3b_free_genomes(cps, ncps);
   return 0;
}
