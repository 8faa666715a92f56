/*
 * flam3_b200_frontend.h — a self-contained front end for the B200 render path (SURVEY.md
 * §8(f) row 4): genome files, the palette library and image files without the reference's
 * parser.c / palettes.c / png.c. Everything works on the reference's public `flam3_genome`
 * (flam3.h), so the results go straight into flam3_render. Implemented in
 * flam3_b200/csrc/f3b_frontend.c; tools/f3b-render.c is a renderer built from nothing else.
 */
#ifndef FLAM3_B200_FRONTEND_H
#define FLAM3_B200_FRONTEND_H

#include "flam3.h"

#ifdef __cplusplus
extern "C" {
#endif

/* flam3_parse_xml2 (flam3.c:1423-1520) with scan_for_flame_nodes / parse_flame_element /
 * parse_xform_xml (parser.c:220-1257): every <flame> of the XML text, in document order.
 * defaults_on = flam3_defaults_on / flam3_defaults_off. Returns NULL (after a message on
 * stderr) on any error. Free with
 * f3b_free_genomes. */
flam3_genome *f3b_parse_flam3(const char *xmltext, int defaults_on, int *ncps);
void f3b_free_genomes(flam3_genome *genomes, int n);

/* flam3_get_palette (palettes.c:152-195): palette `number` of the library file
 * ($flam3_palettes, else ./flam3-palettes.xml) rotated in hue; returns the number, or -1 and
 * an all-white palette. */
int f3b_get_palette(int number, flam3_palette palette, double hue_rotation);

/* clear_cp (flam3.c:1270), flam3_add_xforms + initialize_xforms (flam3.c:1033,
 * variations.c:2401) for one xform (returns its index, or -1), flam3_copy (flam3.c:1171) */
void f3b_genome_init(flam3_genome *cp, int defaults_on);
void f3b_genome_free_xforms(flam3_genome *cp);
int f3b_genome_add_xform(flam3_genome *cp, int is_final);
int f3b_genome_copy(flam3_genome *dest, const flam3_genome *src);

/* flam3_interpolate (flam3.c:717-802): the genome at `time` between the key
 * frames cps[0..ncps) - aligned copies padded to the same xforms (flam3_align,
 * interpolation.c:768-1032), blended with two weights or, on smooth segments, four Catmull-Rom
 * weights (flam3_interpolate_n, interpolation.c:373-708); affine parts in log-polar form; with
 * stagger > 0 the xforms of a linear segment blend one after the other (get_stagger_coef). The
 * result is what the reference's function gives, byte for byte. Returns 0, or 1 out of memory. */
int f3b_interpolate(flam3_genome *cps, int ncps, double time, double stagger, flam3_genome *result);

/* flam3_print_to_string (flam3.c:1628, flam3_print :1674, flam3_print_xform :1899) without edit
 * history or 2.7 compatibility: the genome as XML text, character for character what the
 * reference writes when `version` is its version string ("FLAM3-" + flam3_version(); NULL:
 * "FLAM3-B200"). malloc'd; NULL when out of memory. */
char *f3b_print_flam3(const flam3_genome *cp, const char *version, const char *extra_attributes);

/* binary PPM like flam3-render's format=ppm (flam3-render.c:384-390); 16-bit samples are
 * written big-endian as the format requires */
int f3b_write_ppm(const char *path, const void *image, int width, int height, int nchan, int bytes_per_channel);

#ifdef __cplusplus
}
#endif
#endif
