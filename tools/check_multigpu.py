"""torchrun --nproc-per-node N tools/check_multigpu.py - every rank renders the same frame
(i) alone (all chain groups, whole-frame post-pass) and (ii) together with the other ranks:
chains sharded, strips of the frame reduced out of the peers' memory (CUDA IPC, NVLink), image
rows written into rank 0's image. Rank 0's image must be identical bytes both ways (the
buckets are exact integers, so the sum over ranks is THE histogram); also checks the older
host-callback path (NCCL all-reduce of the uint32 buckets + all-gather of strips)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
os.environ["F3B_DEVICE"] = str(local)
os.environ.pop("F3B_DEVICES", None)
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import flam3_b200 as fb
from flam3_b200.genome_xml import resize, set_flame_attrs
from flam3_b200.multigpu import make_allreduce_exchange, make_strip_gather, setup_peer_strips

core, gpu = fb.load_core(), fb.load_dropin()
f0 = open(os.path.join(ROOT, "tests/golden/test_flame0.flam3")).read()
f1 = open(os.path.join(ROOT, "tests/golden/test_flame1.flam3")).read()
two = "<flames>" + resize(f0, 512, 288, scale=128, quality=40, oversample=2, temporal_samples=30) + \
      resize(f1, 512, 288, scale=128, quality=40, oversample=2, temporal_samples=30).replace('time="0"', 'time="100"', 1) + "</flames>"
ok = True
cases = [("os1", resize(f0, 640, 363, scale=160, quality=50, temporal_samples=2), {}),
         ("os2 earlyclip", resize(f0, 320, 200, scale=80, quality=50, oversample=2), {"earlyclip": 1}),
         ("C3-shaped: os2, two keyframes, time 50, motion blur", two, {"ngenomes": 2, "time": 50.0}),
         ("count plane (linear palette)", set_flame_attrs(resize(f0, 400, 300, scale=100, quality=40, temporal_samples=1), palette_mode="linear"), {}),
         ("C2 full size", resize(f0, 1920, 1080, scale=540, quality=2000, oversample=1, temporal_samples=1), {})]
for name, xml, kw in cases:
    g = gpu.parse(xml)
    sb = {} if "full size" in name else {"sub_batch_size": 200}
    gpu.set_peers(None, None)
    gpu.set_barrier(None)
    gpu.set_shard(0, 1)
    gpu.set_exchange(None)
    gpu.set_strip_gather(None)
    gpu.set_stream_ordered(False)
    alone, st, _ = gpu.render(g, isaac_seed="mg", **sb, **kw)
    # peer path
    setup_peer_strips(gpu, core, rank, world, g, dist=dist, isaac_seed="mg", **sb, **kw)
    peer, st, _ = gpu.render(g, isaac_seed="mg", **sb, **kw)
    peer2, st, _ = gpu.render(g, isaac_seed="mg", **sb, **kw)
    lost, sat = gpu.render_context(core).exactness()
    # host-callback path (all ranks end up with the whole image)
    gpu.set_peers(None, None)
    gpu.set_barrier(None)
    gpu.set_shard(rank, world)
    gpu.set_stream_ordered(True)
    gpu.set_exchange(make_allreduce_exchange(core, stream_ordered=True))
    gpu.set_strip_gather(make_strip_gather(rank, world, core=core, stream_ordered=True))
    cb, st, _ = gpu.render(g, isaac_seed="mg", **sb, **kw)
    same_peer = np.array_equal(alone, peer) and np.array_equal(alone, peer2) if rank == 0 else True
    same_cb = np.array_equal(alone, cb)
    t = torch.tensor([1 if (same_peer and same_cb and lost == 0) else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("%s: N=%d peer strips identical to single GPU: %s; callbacks (all-reduce + all-gather): %s; "
              "staged values at 2^24: %d (mean %.2f)" % (name, world, same_peer, same_cb, lost, alone.mean()), flush=True)
    ok = ok and bool(t.item())
gpu.set_exchange(None)
gpu.set_strip_gather(None)
dist.barrier()
dist.destroy_process_group()
print("check_multigpu rank %d: %s" % (rank, "OK" if ok else "FAILED"), flush=True)
sys.exit(0 if ok else 1)
