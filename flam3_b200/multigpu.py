"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the
histogram exchange. The sample budget shards naturally: every rank runs
the chain groups g with g % nranks == rank into a private histogram, the histograms are
summed, and density estimation + filtering follow.

Only the exchange lives here; the kernels and the sharding rule are in the native library
(f3b_iterate's rank/nranks, f3b_render_set_shard / f3b_render_set_exchange).
"""
import ctypes as C


class _DeviceArray:
    """Minimal __cuda_array_interface__ view of a raw device pointer."""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {
            "shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2,
        }


def device_tensor(ptr, n, typestr="<f4"):
    import torch

    return torch.as_tensor(_DeviceArray(ptr, n, typestr), device="cuda")


def strip_rows(height, rank, nranks):
    """Image rows [row0, row1) a rank post-processes: equal strips of ceil(height/nranks)
    rows, the last one(s) clipped (f3b_render.c: strip_rows_per / strip_row0 / strip_row1)."""
    per = (height + nranks - 1) // nranks
    r0 = min(rank * per, height)
    return r0, min(r0 + per, height), per


def shard_groups(n_groups, rank, nranks):
    """The chain groups a rank runs — the rule f3b_iterate implements on the device
    (f3b_iterate.cu: global group = local * nranks + rank). Pure function, used by the
    CPU-side tests of the sharding logic."""
    return list(range(rank, n_groups, nranks))


def _render_stream(d, ctx):
    """torch view of the drop-in's render stream, so that collectives are ordered with the
    kernels by the stream and not by host synchronisation."""
    import torch

    d.f3b_stream_ptr.restype = C.c_void_p
    d.f3b_stream_ptr.argtypes = [C.c_void_p]
    return torch.cuda.ExternalStream(int(d.f3b_stream_ptr(ctx)))


def make_allreduce_exchange(core, nbuckets=None, dist=None, stream_ordered=False):
    """Returns the callback for Flam3Lib.set_exchange: sums the per-rank histograms in
    place with one all-reduce per plane (the buckets are uint32, summed as int32: exact and
    order independent; no saturation across ranks on this path). stream_ordered=True issues the
    collective on the render stream and returns at once (pair with set_stream_ordered(True))."""
    import torch

    if dist is None:
        import torch.distributed as dist
    d = core._dll
    d.f3b_hist_elems.restype = C.c_size_t
    d.f3b_hist_elems.argtypes = [C.c_void_p]
    cache = {}

    def exchange(ctx_ptr):
        ctx = C.c_void_p(ctx_ptr)
        color = d.f3b_hist_color_ptr(ctx)
        count = d.f3b_hist_count_ptr(ctx)
        n = int(d.f3b_hist_elems(ctx))  # height * row pitch (>= height * width)
        key = (color, count, n)
        if key not in cache:
            cache.clear()
            # the buckets are uint32 (exact integer sums): summed as int32, same bits
            cache[key] = (device_tensor(color, n * 4, "<i4"), device_tensor(count, n, "<i4") if count else None,
                          _render_stream(d, ctx) if stream_ordered else None)
        t, t2, stream = cache[key]
        if stream is not None:
            with torch.cuda.stream(stream):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                if t2 is not None:
                    dist.all_reduce(t2, op=dist.ReduceOp.SUM)
            return
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if t2 is not None:
            dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()

    return exchange


def make_strip_gather(rank, nranks, dist=None, core=None, stream_ordered=False):
    """Returns the callback for Flam3Lib.set_strip_gather: an in-place all-gather of the
    ranks' image strips (rows_per * row_bytes bytes each) so that every rank ends up with
    the whole image."""
    import torch

    if dist is None:
        import torch.distributed as dist
    cache = {}

    def gather(ctx_ptr, image_ptr, row_bytes, rows_per):
        chunk = int(row_bytes) * int(rows_per)
        key = (image_ptr, chunk)
        if key not in cache:
            cache.clear()
            stream = _render_stream(core._dll, C.c_void_p(ctx_ptr)) if (stream_ordered and core is not None) else None
            cache[key] = (device_tensor(image_ptr, chunk * nranks, "|u1"), stream)
        t, stream = cache[key]
        if stream is not None:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(t, t[rank * chunk:(rank + 1) * chunk])
            return
        dist.all_gather_into_tensor(t, t[rank * chunk:(rank + 1) * chunk])
        torch.cuda.synchronize()

    return gather


_opened = {}  # CUDA IPC handle -> device pointer in this process (a handle is opened once)


def _ipc_open(ctx, handle):
    if handle not in _opened:
        _opened[handle] = ctx.ipc_open(handle)
    return _opened[handle]


def setup_peer_strips(gpu, core, rank, nranks, genomes, dist=None, **render_kw):
    """One process per GPU, strips reduced over peer memory (include/flam3_b200.h, "several
    GPUs on one frame"): every rank renders once on its own so that its bucket planes and image
    exist, their CUDA IPC handles go round with an all-gather, and the drop-in is told where
    every rank's buckets and rank 0's image are. From then on a flam3_render call on every
    rank renders ONE frame together: rank r runs its share of the chains, sums the buckets of
    its strip of rows out of the peers' memory over NVLink and writes its image rows into rank
    0's image; only rank 0's `out` receives the picture. The barriers between the steps are
    one-element NCCL all-reduces issued on the render stream (no host synchronisation).
    Must be called again if a later frame is larger (the planes are reallocated)."""
    import torch

    if dist is None:
        import torch.distributed as dist
    gpu.set_peers(None, None)
    gpu.set_exchange(None)
    gpu.set_strip_gather(None)
    gpu.set_shard(rank, nranks)
    gpu.set_device_only(True)
    _, st, _ = gpu.render(genomes, **render_kw)  # allocates buckets, count plane, image
    gpu.set_device_only(False)
    assert st["rv"] == 0, st
    ctx = gpu.render_context(core)
    ctx.sync()
    color, count, _ = ctx.hist_ptrs()
    image = ctx.image_ptr()
    mine = [ctx.ipc_export(color), ctx.ipc_export(count) if count else bytes(64), ctx.ipc_export(image)]
    t = torch.tensor(list(b"".join(mine)), dtype=torch.uint8, device="cuda")
    allh = torch.empty(nranks * t.numel(), dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(allh, t)
    allh = bytes(allh.cpu().tolist())
    peers, image0 = [], None
    for r in range(nranks):
        h = allh[r * 192:(r + 1) * 192]
        if r == rank:
            peers.append((color, count))
            if r == 0:
                image0 = image
            continue
        c = _ipc_open(ctx, h[:64])
        n = _ipc_open(ctx, h[64:128]) if count else None
        peers.append((c, n))
        if r == 0:
            image0 = _ipc_open(ctx, h[128:192])
    gpu.set_peers(peers, image0)
    stream = _render_stream(core._dll, ctx.ptr)
    token = torch.zeros(1, dtype=torch.int32, device="cuda")

    def barrier(_ctx_ptr):
        with torch.cuda.stream(stream):
            dist.all_reduce(token, op=dist.ReduceOp.SUM)

    gpu.set_barrier(barrier)
    gpu.set_stream_ordered(True)
    return peers
