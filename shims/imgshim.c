/*
 * Stand-ins for the reference's png.c / jpeg.c entry points (img.h), because libpng and
 * libjpeg are not installed in this image. Build-environment shim only; not on the hot
 * path. write_png is a real (zlib-based) RGBA PNG encoder so that an unmodified
 * flam3-render / flam3-animate built against the drop-in library produces viewable
 * files; JPEG output and image reading are reported as unsupported.
 */
#include <arpa/inet.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct {
   char *genome;
   char *badvals;
   char *numiters;
   char *rtime;
} flam3_img_comments;

static void put_chunk(FILE *f, const char *tag, const unsigned char *data, size_t n) {
   unsigned char hdr[8];
   uint32_t be = htonl((uint32_t)n);
   memcpy(hdr, &be, 4);
   memcpy(hdr + 4, tag, 4);
   fwrite(hdr, 1, 8, f);
   if (n) fwrite(data, 1, n, f);
   uLong crc = crc32(0L, Z_NULL, 0);
   crc = crc32(crc, hdr + 4, 4);
   if (n) crc = crc32(crc, data, (uInt)n);
   be = htonl((uint32_t)crc);
   fwrite(&be, 1, 4, f);
}

static void put_text(FILE *f, const char *key, const char *val) {
   if (!val) return;
   size_t kl = strlen(key), vl = strlen(val);
   unsigned char *buf = (unsigned char *)malloc(kl + 1 + vl);
   memcpy(buf, key, kl + 1);
   memcpy(buf + kl + 1, val, vl);
   put_chunk(f, "tEXt", buf, kl + 1 + vl);
   free(buf);
}

void write_png(FILE *file, void *image, int width, int height, flam3_img_comments *fpc,
               int bpc) {
   static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
   size_t stride = (size_t)width * 4 * (size_t)bpc;
   size_t rawlen = (stride + 1) * (size_t)height;
   unsigned char *raw = (unsigned char *)malloc(rawlen);
   if (!raw) return;
   for (int y = 0; y < height; y++) {
      unsigned char *dst = raw + (stride + 1) * (size_t)y;
      const unsigned char *src = (const unsigned char *)image + stride * (size_t)y;
      dst[0] = 0; /* filter: none */
      if (bpc == 2) { /* host is little endian on every platform this targets */
         for (size_t i = 0; i < stride; i += 2) {
            dst[1 + i] = src[i + 1];
            dst[2 + i] = src[i];
         }
      } else
         memcpy(dst + 1, src, stride);
   }
   uLongf clen = compressBound((uLong)rawlen);
   unsigned char *comp = (unsigned char *)malloc(clen);
   if (!comp || compress2(comp, &clen, raw, (uLong)rawlen, 6) != Z_OK) {
      fprintf(stderr, "write_png (shim): compression failed\n");
      free(raw);
      free(comp);
      return;
   }
   unsigned char ihdr[13];
   uint32_t be = htonl((uint32_t)width);
   memcpy(ihdr, &be, 4);
   be = htonl((uint32_t)height);
   memcpy(ihdr + 4, &be, 4);
   ihdr[8] = (unsigned char)(8 * bpc);
   ihdr[9] = 6; /* RGBA */
   ihdr[10] = ihdr[11] = ihdr[12] = 0;

   fwrite(sig, 1, 8, file);
   put_chunk(file, "IHDR", ihdr, 13);
   if (fpc) {
      const char *e = getenv("enable_png_comments");
      if (!e || atoi(e) == 1) {
         put_text(file, "flam3_nickname", getenv("nick"));
         put_text(file, "flam3_url", getenv("url"));
         put_text(file, "flam3_id", getenv("id"));
         put_text(file, "flam3_error_rate", fpc->badvals);
         put_text(file, "flam3_samples", fpc->numiters);
         put_text(file, "flam3_time", fpc->rtime);
         put_text(file, "flam3_genome", fpc->genome);
      }
   }
   put_chunk(file, "IDAT", comp, clen);
   put_chunk(file, "IEND", NULL, 0);
   free(raw);
   free(comp);
}

void write_jpeg(FILE *file, unsigned char *image, int width, int height,
                flam3_img_comments *fpc) {
   (void)file; (void)image; (void)width; (void)height; (void)fpc;
   fprintf(stderr, "write_jpeg: libjpeg is not available in this build; use format=png or ppm\n");
}

unsigned char *read_png(FILE *file, int *width, int *height) {
   (void)file; (void)width; (void)height;
   fprintf(stderr, "read_png: not available in this build\n");
   return NULL;
}

unsigned char *read_jpeg(FILE *file, int *width, int *height) {
   (void)file; (void)width; (void)height;
   fprintf(stderr, "read_jpeg: not available in this build\n");
   return NULL;
}
