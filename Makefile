# Top-level build. Everything is built IN-TREE so the artefacts travel with the snapshot:
#
#   make lib      flam3_b200/lib/libflam3_b200.so   CUDA kernels + stage-level C-ABI (sm_100a)
#   make dropin   dropin/_build/libflam3.so         flam3_render drop-in + the reference's host
#                                                   side (parser, interpolation, ...) it plugs into,
#                                                   and the UNMODIFIED CLIs linked against it
#   make oracle   oracle/_ref/*                     unmodified reference (checker / CPU baseline)
#
# `dropin` and `oracle` compile reference sources from $(F3_REF) where they lie; neither
# copies them. On a box without $(F3_REF) only the prebuilt artefacts are used.

F3_REF  ?= /root/reference
NVCC    ?= nvcc
CC      ?= gcc
ARCH     = -gencode arch=compute_100a,code=sm_100a
NVFLAGS  = $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $(NVFLAGS_EXTRA)
CSRC     = flam3_b200/csrc
LIBDIR   = flam3_b200/lib
DROP     = dropin/_build
SHIMS    = shims

all: lib dropin oracle

# ---------------------------------------------------------------- CUDA library
LIB_OBJS = $(LIBDIR)/f3b_iterate.o $(LIBDIR)/f3b_post.o $(LIBDIR)/f3b_api.o $(LIBDIR)/f3b_filters.o

lib: $(LIBDIR)/libflam3_b200.so

$(LIBDIR):
	mkdir -p $(LIBDIR)

$(LIBDIR)/%.o: $(CSRC)/%.cu $(CSRC)/f3b_internal.h $(CSRC)/f3b_pack.h $(CSRC)/f3b_variations.cuh include/flam3_b200.h | $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

# the chaos game tolerates 2-ulp divides / square roots and flushed denormals (the walk is
# statistical; parity is checked by PSNR against the reference and by the per-variation
# single-step test), and they cost 3 instructions instead of ~15
$(LIBDIR)/f3b_iterate.o: NVFLAGS += -prec-div=false -prec-sqrt=false -ftz=true

$(LIBDIR)/f3b_filters.o: $(CSRC)/f3b_filters.c $(CSRC)/f3b_filters.h | $(LIBDIR)
	$(CC) -O2 -std=gnu99 -fPIC -Wall -c $< -o $@

$(LIBDIR)/libflam3_b200.so: $(LIB_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(LIB_OBJS) -lm

# ---------------------------------------------------------------- drop-in libflam3
ifneq ($(wildcard $(F3_REF)/flam3.h),)
REF_CFLAGS = -g -O3 -std=gnu99 -ffast-math -fPIC -w -I$(SHIMS) -I$(F3_REF) \
             -DPACKAGE_DATA_DIR='"$(F3_REF)"' -DGIT_REV='""'
HOST_SRCS  = variations.c filters.c isaac.c interpolation.c palettes.c parser.c
HOST_OBJS  = $(addprefix $(DROP)/host_,$(HOST_SRCS:.c=.o))
GLUE_OBJS  = $(DROP)/f3b_render.o $(DROP)/f3b_pack.o $(DROP)/f3b_isaac.o $(DROP)/f3b_debug.o $(DROP)/f3b_analysis.o \
             $(DROP)/f3b_frontend.o
MISC_OBJS  = $(DROP)/xmlshim.o $(DROP)/imgshim.o
CLIS       = $(DROP)/flam3-render $(DROP)/flam3-animate $(DROP)/flam3-genome

dropin: lib $(DROP)/libflam3.so $(CLIS) $(DROP)/libf3harness.so $(DROP)/f3b-render

$(DROP):
	mkdir -p $(DROP)

# the reference's host side, unmodified, from where it lies
$(DROP)/host_%.o: $(F3_REF)/%.c | $(DROP)
	$(CC) $(REF_CFLAGS) -c $< -o $@

# flam3.c also holds the CPU flam3_render this library replaces: rename it out of the way
$(DROP)/host_flam3.o: $(F3_REF)/flam3.c $(F3_REF)/rect.c | $(DROP)
	$(CC) $(REF_CFLAGS) -Dflam3_render=flam3_render_cpu_unused \
	   -Dflam3_render_memory_required=flam3_render_memory_required_cpu_unused \
	   -Dflam3_estimate_bounding_box=flam3_estimate_bounding_box_cpu_unused \
	   -Dflam3_colorhist=flam3_colorhist_cpu_unused -c $< -o $@

$(DROP)/f3b_%.o: $(CSRC)/f3b_%.c $(CSRC)/f3b_host.h $(CSRC)/f3b_pack.h include/flam3_b200.h | $(DROP)
	$(CC) -O2 -std=gnu99 -fPIC -Wall -Wno-unused-function -I$(SHIMS) -I$(F3_REF) -c $< -o $@

$(DROP)/xmlshim.o: $(SHIMS)/libxml/xmlshim.c | $(DROP)
	$(CC) -O2 -fPIC -std=gnu99 -I$(SHIMS)/libxml -c $< -o $@
$(DROP)/imgshim.o: $(SHIMS)/imgshim.c | $(DROP)
	$(CC) -O2 -fPIC -std=gnu99 -c $< -o $@
$(DROP)/harness.o: tests/harness/f3_harness.c | $(DROP)
	$(CC) -O2 -fPIC -std=gnu99 -w -I$(SHIMS) -I$(F3_REF) -c $< -o $@

# test infrastructure, NOT part of the product: the ctypes harness over flam3.h as its own
# library, linked against the drop-in like any other libflam3 client
$(DROP)/libf3harness.so: $(DROP)/harness.o $(DROP)/libflam3.so
	$(CC) -shared -o $@ $(DROP)/harness.o -L$(DROP) -lflam3 -L$(LIBDIR) -lflam3_b200 \
	   -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)' -lm

$(DROP)/libflam3.so: $(HOST_OBJS) $(DROP)/host_flam3.o $(GLUE_OBJS) $(MISC_OBJS) $(LIBDIR)/libflam3_b200.so
	$(CC) -shared -Wl,-Bsymbolic -o $@ $(HOST_OBJS) $(DROP)/host_flam3.o $(GLUE_OBJS) $(MISC_OBJS) \
	   -L$(LIBDIR) -lflam3_b200 -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)' -lm -lpthread -lz

# a renderer made of this repository's code only: own front end, the
# flam3_render drop-in, libflam3_b200.so - NO reference object (flam3.h is the public header)
STANDALONE_OBJS = $(DROP)/f3b_frontend.o $(DROP)/f3b_standalone.o $(DROP)/f3b_render.o $(DROP)/f3b_pack.o \
                  $(DROP)/f3b_isaac.o $(DROP)/xmlshim.o $(DROP)/imgshim.o
$(DROP)/f3b-render: tools/f3b-render.c $(STANDALONE_OBJS) $(LIBDIR)/libflam3_b200.so
	$(CC) -O2 -std=gnu99 -Wall -I$(SHIMS) -I$(F3_REF) -o $@ $< $(STANDALONE_OBJS) \
	   -L$(LIBDIR) -lflam3_b200 -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)' -lm -lpthread -lz

# the reference's own command-line programs, unmodified, linked against the drop-in
$(DROP)/flam3-%: $(F3_REF)/flam3-%.c $(F3_REF)/docstring.c $(DROP)/libflam3.so
	$(CC) $(REF_CFLAGS) -o $@ $< $(F3_REF)/docstring.c -L$(DROP) -lflam3 \
	   -L$(LIBDIR) -lflam3_b200 -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)' \
	   -lm -lpthread -lz

oracle:
	$(MAKE) -C oracle F3_REF=$(F3_REF)
else
dropin:
	@echo "reference tree $(F3_REF) not present: using prebuilt $(DROP) artefacts"
oracle:
	@echo "reference tree $(F3_REF) not present: using prebuilt oracle/_ref artefacts"
endif

# microbenchmark of the scattered-reduction ceiling (profiles/r01_red_peak_microbench.txt)
tools/micro/red_peak: tools/micro/red_peak.cu
	$(NVCC) $(ARCH) -O3 -o $@ $<

clean:
	rm -rf $(LIBDIR) $(DROP) oracle/_ref tools/micro/red_peak
.PHONY: all lib dropin oracle clean
