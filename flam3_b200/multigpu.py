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


def device_tensor(ptr, n):
    import torch

    return torch.as_tensor(_DeviceArray(ptr, n), device="cuda")


def shard_groups(n_groups, rank, nranks):
    """The chain groups a rank runs — the rule f3b_iterate implements on the device
    (f3b_iterate.cu: global group = local * nranks + rank). Pure function, used by the
    CPU-side tests of the sharding logic."""
    return list(range(rank, n_groups, nranks))


def make_allreduce_exchange(core, nbuckets, dist=None):
    """Returns the callback for Flam3Lib.set_exchange: sums the per-rank histograms in
    place with one all-reduce per plane (float32; the adds are integer valued, so the sum
    is exact below 2^24 per channel and order independent)."""
    import torch

    if dist is None:
        import torch.distributed as dist
    d = core._dll

    def exchange(ctx_ptr):
        ctx = C.c_void_p(ctx_ptr)
        color = d.f3b_hist_color_ptr(ctx)
        count = d.f3b_hist_count_ptr(ctx)
        t = device_tensor(color, nbuckets * 4)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if count:
            t2 = device_tensor(count, nbuckets)
            dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()

    return exchange
