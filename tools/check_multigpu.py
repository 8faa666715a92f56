"""torchrun --nproc-per-node N tools/check_multigpu.py — every rank renders the same frame
(i) alone (all chain groups, whole-frame post-pass) and (ii) sharded over the N ranks with the
NCCL histogram all-reduce and the row-strip post-pass; the images must be identical bytes
on every rank (bucket adds are integer valued, so the summed histogram is exact)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
os.environ["F3B_DEVICE"] = str(local)
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import flam3_b200 as fb
from flam3_b200.genome_xml import resize
from flam3_b200.multigpu import make_allreduce_exchange, make_strip_gather

core, gpu = fb.load_core(), fb.load_dropin()
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
ok = True
for name, xml, kw in [("os1", resize(f0, 640, 363, scale=160, quality=50, temporal_samples=2), {}),
                      ("os2 earlyclip", resize(f0, 320, 200, scale=80, quality=50, oversample=2), {"earlyclip": 1})]:
    g = gpu.parse(xml)
    geo = gpu.buffer_geometry(g)
    gpu.set_shard(0, 1)
    gpu.set_exchange(None)
    gpu.set_strip_gather(None)
    alone, st, _ = gpu.render(g, isaac_seed="mg", sub_batch_size=200)
    gpu.set_shard(rank, world)
    gpu.set_stream_ordered(False)
    gpu.set_exchange(make_allreduce_exchange(core))
    whole, st, _ = gpu.render(g, isaac_seed="mg", sub_batch_size=200)
    gpu.set_strip_gather(make_strip_gather(rank, world))
    strips, st, _ = gpu.render(g, isaac_seed="mg", sub_batch_size=200)
    # the same with the collectives issued on the render stream (no host synchronisation)
    gpu.set_stream_ordered(True)
    gpu.set_exchange(make_allreduce_exchange(core, stream_ordered=True))
    gpu.set_strip_gather(make_strip_gather(rank, world, core=core, stream_ordered=True))
    ordered, st, _ = gpu.render(g, isaac_seed="mg", sub_batch_size=200)
    ordered2, st, _ = gpu.render(g, isaac_seed="mg", sub_batch_size=200)
    gpu.set_stream_ordered(False)
    same = (np.array_equal(alone, whole) and np.array_equal(alone, strips) and
            np.array_equal(alone, ordered) and np.array_equal(alone, ordered2))
    t = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("%s: N=%d sharded+allreduce %s, +strips %s, stream-ordered %s (mean %.2f)" % (
            name, world, np.array_equal(alone, whole), np.array_equal(alone, strips),
            np.array_equal(alone, ordered) and np.array_equal(alone, ordered2), alone.mean()), flush=True)
    ok = ok and bool(t.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
