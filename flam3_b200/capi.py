"""ctypes bindings.

``CoreLib``  — include/flam3_b200.h (stage-level C-ABI of libflam3_b200.so).
``Flam3Lib`` — the flam3.h API of a libflam3 build (the drop-in by default) through the
               small C harness tests/harness/f3_harness.c, the way flam3-render.c and
               flam3-animate.c use it: parse XML -> fill a flam3_frame -> flam3_render.

The same ``Flam3Lib`` class can be pointed at any library exporting that harness; the test
suite points a second instance at the unmodified reference build to compare against.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def repo_root():
    return os.path.dirname(_HERE)


def core_lib_path():
    return os.path.join(_HERE, "lib", "libflam3_b200.so")


def dropin_lib_path():
    return os.path.join(repo_root(), "dropin", "_build", "libflam3.so")


def dropin_harness_path():
    """The test harness over flam3.h (tests/harness/f3_harness.c) built as its own library
    against the drop-in libflam3.so — the product library itself holds no test code."""
    # (F3B_HARNESS_LIB: another build of the same pair of libraries, e.g. one made with
    # -fsanitize=address for the host-side tests)
    return os.environ.get("F3B_HARNESS_LIB") or os.path.join(repo_root(), "dropin", "_build", "libf3harness.so")


class RenderArgs(C.Structure):
    """Mirror of h_render_args (tests/harness/f3_harness.c)."""

    _fields_ = [
        ("bits", C.c_int),
        ("bytes_per_channel", C.c_int),
        ("earlyclip", C.c_int),
        ("nthreads", C.c_int),
        ("sub_batch_size", C.c_int),
        ("verbose", C.c_int),
        ("field", C.c_int),
        ("nchan", C.c_int),
        ("transp", C.c_int),
        ("first_genome", C.c_int),
        ("ngenomes", C.c_int),
        ("time", C.c_double),
        ("pixel_aspect_ratio", C.c_double),
        ("isaac_seed", C.c_char * 128),
    ]


def make_args(**kw):
    a = RenderArgs()
    a.bits = 33
    a.bytes_per_channel = 1
    a.earlyclip = 0
    a.nthreads = 1
    a.sub_batch_size = 10000
    a.verbose = 0
    a.field = 0
    a.nchan = 3
    a.transp = 0
    a.first_genome = 0
    a.ngenomes = 1
    a.time = 0.0
    a.pixel_aspect_ratio = 1.0
    a.isaac_seed = b"flam3b200"
    for k, v in kw.items():
        if k == "isaac_seed" and isinstance(v, str):
            v = v.encode()
        if not hasattr(a, k):
            raise AttributeError(k)
        setattr(a, k, v)
    return a


class Genomes:
    """Owning handle of a parsed genome array inside one library."""

    def __init__(self, lib, ptr, n):
        self.lib, self.ptr, self.n = lib, ptr, n

    def dims(self, idx=0):
        w, h, o = C.c_int(), C.c_int(), C.c_int()
        self.lib._dll.h_genome_dims(self.ptr, idx, C.byref(w), C.byref(h), C.byref(o))
        return w.value, h.value, o.value

    def close(self):
        if self.ptr:
            self.lib._dll.h_free_genomes(self.ptr, self.n)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Flam3Lib:
    """A libflam3 build (flam3.h API + harness). Default: the B200 drop-in."""

    def __init__(self, path=None):
        path = path or dropin_harness_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "native library %s is missing: run `make` (or __graft_entry__.build()) first; "
                "there is no Python fallback" % path)
        self.path = path
        # RTLD_LOCAL + -Bsymbolic at link time: two libflam3 builds can coexist in-process
        self._dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        d = self._dll
        d.h_parse.restype = C.c_void_p
        d.h_parse.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        d.h_free_genomes.argtypes = [C.c_void_p, C.c_int]
        d.h_genome_dims.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
        d.h_render.restype = C.c_int
        d.h_render.argtypes = [C.c_void_p, C.POINTER(RenderArgs), C.c_void_p,
                               C.POINTER(C.c_double), C.c_void_p]
        d.h_memory_required.restype = C.c_double
        d.h_memory_required.argtypes = [C.c_void_p, C.c_int, C.c_int]
        d.h_buffer_geometry.restype = C.c_int
        d.h_buffer_geometry.argtypes = [C.c_void_p, C.POINTER(RenderArgs), C.POINTER(C.c_int)]
        d.h_print_genome.restype = C.c_void_p
        d.h_print_genome.argtypes = [C.c_void_p, C.c_int]
        d.h_abi_sizes.argtypes = [C.POINTER(C.c_int)]

    # ---- drop-in only: post-pass from a supplied histogram, and the render context ----
    def inject_histogram(self, buckets5):
        """Make the following flam3_render calls skip the chaos game and use `buckets5`
        (uint32[h, w, 5], the reference's bits=33 bucket layout). None switches back."""
        f = self._dll.f3b_render_inject_histogram
        f.argtypes = [C.c_void_p, C.c_size_t]
        f.restype = None
        if buckets5 is None:
            self._inject_keep = None
            f(None, 0)
            return
        b = np.ascontiguousarray(buckets5, dtype=np.uint32)
        self._inject_keep = b
        f(b.ctypes.data_as(C.c_void_p), b.size // 5)

    def preload_histogram(self, buckets5):
        """Like inject_histogram, but the chaos game still runs and adds on top."""
        f = self._dll.f3b_render_preload_histogram
        f.argtypes = [C.c_void_p, C.c_size_t]
        f.restype = None
        if buckets5 is None:
            self._inject_keep = None
            f(None, 0)
            return
        b = np.ascontiguousarray(buckets5, dtype=np.uint32)
        self._inject_keep = b
        f(b.ctypes.data_as(C.c_void_p), b.size // 5)

    def set_device_only(self, on):
        self._dll.f3b_render_set_device_only(int(bool(on)))

    def set_shard(self, rank, nranks):
        self._dll.f3b_render_set_shard(int(rank), int(nranks))

    def set_exchange(self, fn):
        """fn(ctx_ptr) is called after every batch's chaos game (histogram exchange hook)."""
        proto = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
        f = self._dll.f3b_render_set_exchange
        f.argtypes = [C.c_void_p, C.c_void_p]
        f.restype = None
        if fn is None:
            self._exchange_keep = None
            f(None, None)
            return
        cb = proto(lambda ctx, user: fn(ctx))
        self._exchange_keep = cb
        f(C.cast(cb, C.c_void_p), None)

    def set_stream_ordered(self, on):
        """Callbacks enqueue on the render stream instead of being bracketed by host syncs."""
        self._dll.f3b_render_set_stream_ordered(int(bool(on)))

    def set_strip_gather(self, fn):
        """fn(ctx_ptr, image_ptr, row_bytes, rows_per) assembles the per-rank image strips
        (include/flam3_b200_dropin.h: f3b_render_set_strip_gather). None = whole-frame
        post-pass on every rank."""
        proto = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)
        f = self._dll.f3b_render_set_strip_gather
        f.argtypes = [C.c_void_p, C.c_void_p]
        f.restype = None
        if fn is None:
            self._gather_keep = None
            f(None, None)
            return
        cb = proto(lambda ctx, image, row_bytes, rows_per, user: fn(ctx, image, row_bytes, rows_per))
        self._gather_keep = cb
        f(C.cast(cb, C.c_void_p), None)

    def set_peers(self, peers, image0):
        """Ranks in separate processes: peers = [(color_ptr, count_ptr or None)] per rank, as
        device pointers valid in THIS process (CUDA IPC), image0 = rank 0's device image.
        None switches the peer path off (include/flam3_b200_dropin.h)."""
        f = self._dll.f3b_render_set_peers
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        f.restype = None
        if not peers:
            f(0, None, None)
            return
        arr = (PeerHist * len(peers))()
        for i, (c, n) in enumerate(peers):
            arr[i].color, arr[i].count = c, n
        f(len(peers), C.cast(arr, C.c_void_p), image0)

    def set_barrier(self, fn):
        """fn(ctx_ptr) must enqueue a barrier over all ranks on the render stream."""
        proto = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
        f = self._dll.f3b_render_set_barrier
        f.argtypes = [C.c_void_p, C.c_void_p]
        f.restype = None
        if fn is None:
            self._barrier_keep = None
            f(None, None)
            return
        cb = proto(lambda ctx, user: fn(ctx))
        self._barrier_keep = cb
        f(C.cast(cb, C.c_void_p), None)

    def render_context_at(self, core, i):
        """The i-th GPU context of this process (F3B_DEVICES), or None."""
        f = self._dll.f3b_render_context_at
        f.restype = C.c_void_p
        f.argtypes = [C.c_int]
        p = f(i)
        return Context(core, C.c_void_p(p)) if p else None

    def render_context(self, core):
        """The f3b_ctx the drop-in renders with, wrapped for stage stats / buffer access."""
        f = self._dll.f3b_render_context
        f.restype = C.c_void_p
        f.argtypes = []
        p = f()
        if not p:
            raise RuntimeError("drop-in has no GPU context")
        return Context(core, C.c_void_p(p))

    def parse(self, xml):
        if isinstance(xml, str):
            xml = xml.encode()
        n = C.c_int(0)
        p = self._dll.h_parse(xml, C.byref(n))
        if not p or n.value <= 0:
            raise ValueError("flam3_parse_xml2 found no genomes")
        return Genomes(self, p, n.value)

    def parse_own(self, xml):
        """Parse with the repository's own front end (f3b_parse_flam3; drop-in library only)."""
        if isinstance(xml, str):
            xml = xml.encode()
        f = self._dll.h_parse_own
        f.restype = C.c_void_p
        f.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        n = C.c_int(0)
        p = f(xml, C.byref(n))
        if not p or n.value <= 0:
            raise ValueError("f3b_parse_flam3 found no genomes")
        return Genomes(self, p, n.value)

    def abi_sizes(self):
        out = (C.c_int * 4)()
        self._dll.h_abi_sizes(out)
        return dict(randctx=out[0], flam3_xform=out[1], flam3_genome=out[2], flam3_frame=out[3])

    def buffer_geometry(self, genomes, **kw):
        a = make_args(**kw)
        dims = (C.c_int * 6)()
        if self._dll.h_buffer_geometry(genomes.ptr, C.byref(a), dims):
            raise RuntimeError("buffer geometry failed")
        return dict(width=dims[0], height=dims[1], filter_width=dims[2], gutter=dims[3],
                    de_offset=dims[4], oversample=dims[5])

    def memory_required(self, genomes, idx=0, bits=33):
        return self._dll.h_memory_required(genomes.ptr, idx, bits)

    def print_genome(self, genomes, idx=0):
        p = self._dll.h_print_genome(genomes.ptr, idx)
        s = C.string_at(p).decode()
        return s

    def render_with_progress(self, genomes, abort_at=0, **kw):
        """flam3_render with a progress callback installed (flam3_frame.progress): returns
        (image, stats, dict(calls, max_percent, last_eta))."""
        a = make_args(**kw)
        w, h, _ = genomes.dims(a.first_genome)
        out = np.zeros((h, w, a.nchan), dtype=np.uint8 if a.bytes_per_channel == 1 else np.uint16)
        stats, prog = (C.c_double * 3)(), (C.c_double * 3)()
        f = self._dll.h_render_progress
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.POINTER(RenderArgs), C.c_void_p, C.POINTER(C.c_double), C.c_int,
                      C.POINTER(C.c_double)]
        rv = f(genomes.ptr, C.byref(a), out.ctypes.data_as(C.c_void_p), stats, abort_at, prog)
        st = dict(rv=rv, badvals=stats[0], num_iters=stats[1], render_seconds=stats[2])
        return out, st, dict(calls=int(prog[0]), max_percent=prog[1], last_eta=prog[2])

    def bounding_box(self, genomes, idx=0, eps=0.01, nsamples=100000, seed="flam3b200"):
        """flam3_estimate_bounding_box: (rv, (xmin, ymin, xmax, ymax))."""
        f = self._dll.h_bounding_box
        f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_char_p, C.POINTER(C.c_double)]
        out = (C.c_double * 4)()
        rv = f(genomes.ptr, idx, eps, nsamples, seed.encode(), out)
        return rv, tuple(out)

    def colorhist(self, genomes, idx=0, num_batches=10, seed="flam3b200"):
        """flam3_colorhist: (rv, hist[256])."""
        f = self._dll.h_colorhist
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_double)]
        out = (C.c_double * 256)()
        rv = f(genomes.ptr, idx, num_batches, seed.encode(), out)
        return rv, np.array(out[:])

    def render(self, genomes, out=None, **kw):
        """flam3_render of genomes[first_genome:first_genome+ngenomes].

        Returns (image ndarray (h, w, nchan) u8/u16, stats dict, rc bytes)."""
        a = make_args(**kw)
        w, h, _ = genomes.dims(a.first_genome)
        dt = np.uint8 if a.bytes_per_channel == 1 else np.uint16
        if out is None:
            out = np.zeros((h, w, a.nchan), dtype=dt)
        assert out.flags["C_CONTIGUOUS"] and out.nbytes >= h * w * a.nchan * a.bytes_per_channel
        stats = (C.c_double * 3)()
        rc = C.create_string_buffer(288)
        rv = self._dll.h_render(genomes.ptr, C.byref(a), out.ctypes.data_as(C.c_void_p), stats, rc)
        st = dict(rv=rv, badvals=stats[0], num_iters=stats[1], render_seconds=stats[2])
        return out, st, rc.raw


class PeerHist(C.Structure):
    """f3b_peer_hist (include/flam3_b200.h)."""

    _fields_ = [("color", C.c_void_p), ("count", C.c_void_p)]


class F3Job(C.Structure):
    _fields_ = [("blob_offset", C.c_uint64), ("n_iters", C.c_uint64), ("n_chains", C.c_uint32),
                ("fuse", C.c_uint32), ("first_group", C.c_uint32), ("sample_id", C.c_uint32)]


class DensityDesc(C.Structure):
    _fields_ = [("k1", C.c_double), ("k2", C.c_double), ("oversample", C.c_int),
                ("int_accum", C.c_int), ("n_kernels", C.c_int), ("kernel_size", C.c_int),
                ("max_filtered_counts", C.c_int), ("max_filter_index", C.c_int),
                ("curve", C.c_double), ("coefs", C.POINTER(C.c_double)),
                ("widths", C.POINTER(C.c_double))]


class FinalDesc(C.Structure):
    _fields_ = [("image_width", C.c_int), ("image_height", C.c_int), ("oversample", C.c_int),
                ("de_offset", C.c_int), ("filter_width", C.c_int),
                ("filter", C.POINTER(C.c_double)), ("earlyclip", C.c_int), ("nchan", C.c_int),
                ("transp", C.c_int), ("bytes_per_channel", C.c_int), ("gamma", C.c_double),
                ("vibrancy", C.c_double), ("background", C.c_double * 3),
                ("highlight_power", C.c_double), ("gam_lin_thresh", C.c_double)]


class DeFilters(C.Structure):
    """f3b_de_filters (flam3_b200/csrc/f3b_filters.h)."""

    _fields_ = [("max_filtered_counts", C.c_int), ("max_filter_index", C.c_int),
                ("kernel_size", C.c_int), ("n_kernels", C.c_int),
                ("widths", C.POINTER(C.c_double)), ("coefs", C.POINTER(C.c_double))]


# every symbol include/flam3_b200.h declares (checked by the CPU test-suite)
CORE_SYMBOLS = [
    "f3b_abi_version", "f3b_ctx_create", "f3b_ctx_destroy", "f3b_last_error", "f3b_device_info",
    "f3b_sync", "f3b_frame_begin", "f3b_hist_clear", "f3b_accum_clear", "f3b_plan_chains",
    "f3b_iterate", "f3b_pinned_alloc", "f3b_pinned_free", "f3b_inputs_consumed", "f3b_badvals", "f3b_exactness", "f3b_density", "f3b_density_rows", "f3b_finish",
    "f3b_finish_device", "f3b_finish_device_rows", "f3b_image_download", "f3b_hist_upload_u32", "f3b_hist_download_u32", "f3b_accum_upload",
    "f3b_accum_download", "f3b_hist_color_ptr", "f3b_hist_count_ptr", "f3b_hist_elems", "f3b_accum_ptr",
    "f3b_stream_ptr", "f3b_stage_stats", "f3b_stats_reset", "f3b_xfer_stats", "f3b_timer_start",
    "f3b_timer_stop", "f3b_peer_enable", "f3b_hist_reduce_rows", "f3b_group_barrier",
    "f3b_image_reserve", "f3b_image_ptr", "f3b_finish_rows_into", "f3b_ipc_export", "f3b_ipc_open",
    "f3b_ipc_close", "f3b_device_index", "f3b_orbits", "f3b_debug_counters", "f3b_debug_stage_layout",
]


class CoreLib:
    """include/flam3_b200.h over ctypes."""

    def __init__(self, path=None):
        path = path or core_lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "CUDA library %s is missing: run `make lib`; there is no fallback" % path)
        self.path = path
        d = self._dll = C.CDLL(path, mode=C.RTLD_GLOBAL)
        vp, ci = C.c_void_p, C.c_int
        d.f3b_abi_version.restype = ci
        d.f3b_ctx_create.argtypes = [ci, C.POINTER(vp)]
        d.f3b_ctx_destroy.argtypes = [vp]
        d.f3b_last_error.restype = C.c_char_p
        d.f3b_last_error.argtypes = [vp]
        d.f3b_device_info.argtypes = [vp, C.POINTER(ci)]
        d.f3b_sync.argtypes = [vp]
        d.f3b_frame_begin.argtypes = [vp, ci, ci, ci, ci]
        d.f3b_hist_clear.argtypes = [vp]
        d.f3b_accum_clear.argtypes = [vp]
        d.f3b_plan_chains.restype = C.c_uint32
        d.f3b_plan_chains.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        d.f3b_iterate.argtypes = [vp, vp, C.c_size_t, C.POINTER(F3Job), ci, C.c_uint64, ci,
                                  C.c_uint32, C.c_uint32, ci, ci]
        d.f3b_badvals.argtypes = [vp, C.POINTER(C.c_double)]
        d.f3b_density.argtypes = [vp, C.POINTER(DensityDesc)]
        d.f3b_density_rows.argtypes = [vp, C.POINTER(DensityDesc), ci, ci]
        d.f3b_finish.argtypes = [vp, C.POINTER(FinalDesc), vp, C.c_size_t]
        d.f3b_finish_device.argtypes = [vp, C.POINTER(FinalDesc), C.POINTER(vp)]
        d.f3b_hist_upload_u32.argtypes = [vp, vp]
        d.f3b_hist_download_u32.argtypes = [vp, vp]
        d.f3b_accum_upload.argtypes = [vp, vp]
        d.f3b_accum_download.argtypes = [vp, vp]
        for n in ("f3b_hist_color_ptr", "f3b_hist_count_ptr", "f3b_hist_elems", "f3b_accum_ptr", "f3b_stream_ptr"):
            getattr(d, n).restype = vp
            getattr(d, n).argtypes = [vp]
        d.f3b_stage_stats.argtypes = [vp, ci, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        d.f3b_stats_reset.argtypes = [vp]
        # host-side filter construction lives in the same library
        d.f3b_spatial_filter_create.restype = ci
        d.f3b_spatial_filter_create.argtypes = [ci, ci, C.c_double, C.c_double, ci,
                                                C.POINTER(C.POINTER(C.c_double))]
        d.f3b_temporal_filter_create.restype = C.c_double
        d.f3b_temporal_filter_create.argtypes = [ci, ci, C.c_double, C.c_double,
                                                 C.POINTER(C.POINTER(C.c_double)),
                                                 C.POINTER(C.POINTER(C.c_double))]
        d.f3b_de_filters_create.restype = ci
        d.f3b_de_filters_create.argtypes = [C.c_double, C.c_double, C.c_double, ci,
                                            C.POINTER(DeFilters)]
        d.f3b_de_filters_free.argtypes = [C.POINTER(DeFilters)]
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [vp]

    # ---- host-side filters -------------------------------------------------------
    def spatial_filter(self, select, oversample, radius, aspect=1.0, field=0):
        p = C.POINTER(C.c_double)()
        w = self._dll.f3b_spatial_filter_create(select, oversample, radius, aspect, field,
                                                C.byref(p))
        if w < 0:
            return w, None
        arr = np.ctypeslib.as_array(p, shape=(w * w,)).copy()
        self._libc.free(C.cast(p, C.c_void_p))
        return w, arr

    def temporal_filter(self, n, ftype, fexp, fwidth):
        pw, pd = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        s = self._dll.f3b_temporal_filter_create(n, ftype, fexp, fwidth, C.byref(pw), C.byref(pd))
        w = np.ctypeslib.as_array(pw, shape=(n,)).copy()
        d = np.ctypeslib.as_array(pd, shape=(n,)).copy()
        self._libc.free(C.cast(pw, C.c_void_p))
        self._libc.free(C.cast(pd, C.c_void_p))
        return s, w, d

    def de_filters(self, max_rad, min_rad, curve, oversample):
        de = DeFilters()
        rv = self._dll.f3b_de_filters_create(max_rad, min_rad, curve, oversample, C.byref(de))
        if rv:
            return None
        out = dict(max_filtered_counts=de.max_filtered_counts,
                   max_filter_index=de.max_filter_index, kernel_size=de.kernel_size,
                   n_kernels=de.n_kernels,
                   widths=np.ctypeslib.as_array(de.widths, shape=(de.n_kernels,)).copy(),
                   coefs=np.ctypeslib.as_array(
                       de.coefs, shape=(de.n_kernels * de.kernel_size,)).copy())
        self._dll.f3b_de_filters_free(C.byref(de))
        return out

    # ---- GPU context ---------------------------------------------------------------
    def create_context(self, device=0):
        ctx = C.c_void_p()
        if self._dll.f3b_ctx_create(device, C.byref(ctx)) or not ctx:
            raise RuntimeError("f3b_ctx_create(%d) failed: no usable sm_100 GPU" % device)
        return Context(self, ctx)


class Context:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr
        self.width = self.height = 0

    def _ck(self, rv, what):
        if rv:
            raise RuntimeError("%s: %s" % (what, self.lib._dll.f3b_last_error(self.ptr).decode()))

    def close(self):
        if self.ptr:
            self.lib._dll.f3b_ctx_destroy(self.ptr)
            self.ptr = None

    def device_info(self):
        info = (C.c_int * 3)()
        self._ck(self.lib._dll.f3b_device_info(self.ptr, info), "device_info")
        return dict(sms=info[0], blocks_per_sm=info[1], threads=info[2])

    def frame_begin(self, width, height, count_channel=True, bits=33):
        self._ck(self.lib._dll.f3b_frame_begin(self.ptr, width, height, int(count_channel), bits),
                 "frame_begin")
        self.width, self.height = width, height

    def hist_upload_u32(self, buckets5):
        b = np.ascontiguousarray(buckets5, dtype=np.uint32)
        assert b.size == self.width * self.height * 5
        self._ck(self.lib._dll.f3b_hist_upload_u32(self.ptr, b.ctypes.data_as(C.c_void_p)),
                 "hist_upload")

    def hist_download_u32(self):
        b = np.zeros((self.height, self.width, 5), dtype=np.uint32)
        self._ck(self.lib._dll.f3b_hist_download_u32(self.ptr, b.ctypes.data_as(C.c_void_p)),
                 "hist_download")
        return b

    def accum_download(self):
        a = np.zeros((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib._dll.f3b_accum_download(self.ptr, a.ctypes.data_as(C.c_void_p)),
                 "accum_download")
        return a

    def accum_upload(self, acc):
        a = np.ascontiguousarray(acc, dtype=np.float32)
        assert a.size == self.width * self.height * 4
        self._ck(self.lib._dll.f3b_accum_upload(self.ptr, a.ctypes.data_as(C.c_void_p)),
                 "accum_upload")

    def density(self, k1, k2, oversample, de=None, curve=0.4, int_accum=False, rows=None):
        d = DensityDesc()
        d.k1, d.k2, d.oversample, d.int_accum = k1, k2, oversample, int(int_accum)
        keep = None
        if de is not None:
            coefs = np.ascontiguousarray(de["coefs"], dtype=np.float64)
            widths = np.ascontiguousarray(de["widths"], dtype=np.float64)
            keep = (coefs, widths)
            d.n_kernels = de["n_kernels"]
            d.kernel_size = de["kernel_size"]
            d.max_filtered_counts = de["max_filtered_counts"]
            d.max_filter_index = de["max_filter_index"]
            d.curve = curve
            d.coefs = coefs.ctypes.data_as(C.POINTER(C.c_double))
            d.widths = widths.ctypes.data_as(C.POINTER(C.c_double))
        if rows is None:
            self._ck(self.lib._dll.f3b_density(self.ptr, C.byref(d)), "density")
        else:
            self._ck(self.lib._dll.f3b_density_rows(self.ptr, C.byref(d), rows[0], rows[1]),
                     "density_rows")
        del keep

    def finish(self, image_width, image_height, oversample, de_offset, filter_width, filt,
               gamma, vibrancy, background, highlight_power, gam_lin_thresh, earlyclip=0,
               nchan=3, transp=0, bytes_per_channel=1):
        f = FinalDesc()
        filt = np.ascontiguousarray(filt, dtype=np.float64)
        f.image_width, f.image_height = image_width, image_height
        f.oversample, f.de_offset, f.filter_width = oversample, de_offset, filter_width
        f.filter = filt.ctypes.data_as(C.POINTER(C.c_double))
        f.earlyclip, f.nchan, f.transp, f.bytes_per_channel = earlyclip, nchan, transp, bytes_per_channel
        f.gamma, f.vibrancy = gamma, vibrancy
        for i in range(3):
            f.background[i] = background[i]
        f.highlight_power, f.gam_lin_thresh = highlight_power, gam_lin_thresh
        dt = np.uint8 if bytes_per_channel == 1 else np.uint16
        out = np.zeros((image_height, image_width, nchan), dtype=dt)
        self._ck(self.lib._dll.f3b_finish(self.ptr, C.byref(f), out.ctypes.data_as(C.c_void_p),
                                          image_width * nchan * bytes_per_channel), "finish")
        return out

    def exactness(self):
        """(lost, saturated) of the last frame: staged sums that reached 2^24 before they were
        moved into the integer buckets, and uint32 sums pinned at UINT_MAX."""
        a, b = C.c_uint64(), C.c_uint64()
        f = self.lib._dll.f3b_exactness
        f.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        self._ck(f(self.ptr, C.byref(a), C.byref(b)), "exactness")
        return a.value, b.value

    def stage_stats(self, stage):
        ms, n = C.c_double(), C.c_uint64()
        self._ck(self.lib._dll.f3b_stage_stats(self.ptr, stage, C.byref(ms), C.byref(n)),
                 "stage_stats")
        return ms.value, n.value

    def stats_reset(self):
        self.lib._dll.f3b_stats_reset(self.ptr)

    def xfer_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        f = self.lib._dll.f3b_xfer_stats
        f.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        self._ck(f(self.ptr, C.byref(a), C.byref(b)), "xfer_stats")
        return a.value, b.value

    def timer_start(self):
        f = self.lib._dll.f3b_timer_start
        f.argtypes = [C.c_void_p]
        self._ck(f(self.ptr), "timer_start")

    def timer_stop(self):
        ms = C.c_double()
        f = self.lib._dll.f3b_timer_stop
        f.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self._ck(f(self.ptr, C.byref(ms)), "timer_stop")
        return ms.value

    def sync(self):
        self._ck(self.lib._dll.f3b_sync(self.ptr), "sync")

    def image_ptr(self):
        f = self.lib._dll.f3b_image_ptr
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p]
        return f(self.ptr)

    def ipc_export(self, dev_ptr):
        buf = (C.c_ubyte * 64)()
        f = self.lib._dll.f3b_ipc_export
        f.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_ubyte)]
        self._ck(f(self.ptr, dev_ptr, buf), "ipc_export")
        return bytes(buf)

    def ipc_open(self, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        f = self.lib._dll.f3b_ipc_open
        f.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]
        self._ck(f(self.ptr, buf, C.byref(out)), "ipc_open")
        return out.value

    def hist_ptrs(self):
        d = self.lib._dll
        return d.f3b_hist_color_ptr(self.ptr), d.f3b_hist_count_ptr(self.ptr), d.f3b_accum_ptr(self.ptr)


def load_core(path=None):
    return CoreLib(path)


def load_dropin(path=None):
    return Flam3Lib(path)
