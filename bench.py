#!/usr/bin/env python
"""bench.py — headline benchmark of the flam3 render hot path on B200.

Metric (BASELINE.json): chaos-game iterations/sec (and the 1080p q2000 frame time it
implies). A "step" is one full render of the workload frame: chaos game + histogram,
log-scale + density estimation, spatial filter + tone map.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload at every N: config C2 of SURVEY.md §8(d) — test.flam3's first flame at 1920x1080,
quality 2000, oversample 1, estimator_radius 9 (4.147e9 stored iterations per frame). With
N > 1 (one process per GPU) the frame's chains are split over the ranks (strong scaling);
rank r then sums, over NVLink and straight out of the peers' memory, all ranks' buckets of
the rows of image strip r (+ halo), post-processes that strip and writes its rows into rank
0's image (include/flam3_b200.h "several GPUs on one frame"; CUDA IPC handles and the
stream-ordered NCCL barriers are set up by flam3_b200/multigpu.py). --workload c3 / c4 run
the other BASELINE.json configs (4K oversample-2 motion blur; animation frames).

 value : whole-job iterations/s from CUDA events on the render stream around the K
         steps, images left on the device (inputs are a few KB of packed genome).
 e2e   : the same K steps through the reference-facing call flam3_render(flam3_frame*,
         out, ...) with a HOST output buffer: packed genome H2D and image D2H inside the
         timed region.
 --impl reference : the UNMODIFIED reference (oracle/_ref, built from /root/reference) on
         all host cores, on a bounded sample of the same workload (same genome and
         size, reduced quality), rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chaos_game_iterations_per_sec"
UNIT = "iters/s"

# BASELINE.json configs. c2 is the headline line; c3 / c4 are extra
# lines for the configs the metric is not quoted on (python bench.py --workload c3|c4).
WORKLOADS = {
    "c2": dict(name="C2: test.flam3 flame 0 @1920x1080 q2000 os1 estimator_radius 9 (flam3_render, bits=33)",
               w=1920, h=1080, quality=2000.0, scale=540.0, oversample=1, temporal_samples=1,
               keyframes=1, render_kw={}),
    "c3": dict(name="C3: test.flam3 flames 0->1 at time 50 @3840x2160 q5000 os2, 1000 temporal samples "
                    "(motion blur), estimator_radius 9 (flam3_render, bits=33)",
               w=3840, h=2160, quality=5000.0, scale=1080.0, oversample=2, temporal_samples=1000,
               keyframes=2, render_kw={"ngenomes": 2, "time": 50.0}),
    "c4": dict(name="C4: flam3-animate frames 0..99 between test.flam3's flames @640x480 q10, 1000 temporal "
                    "samples each, nthreads=16 (flam3_render per frame, bits=33)",
               w=640, h=480, quality=10.0, scale=240.0, oversample=1, temporal_samples=1000,
               keyframes=2, render_kw={"ngenomes": 2, "nthreads": 16}),
}
WL = WORKLOADS["c2"]


def workload_xml(quality=None, wl=None):
    from flam3_b200.genome_xml import resize

    wl = wl or WL
    q = wl["quality"] if quality is None else quality
    f0 = open(os.path.join(ROOT, "tests", "golden", "test_flame0.flam3")).read()
    # flam3-render forces ntemporal_samples = 1 for stills (flam3-render.c:235)
    a = resize(f0, wl["w"], wl["h"], scale=wl["scale"], quality=q, oversample=wl["oversample"],
               estimator_radius=9, temporal_samples=wl["temporal_samples"])
    if wl["keyframes"] == 1:
        return a
    f1 = open(os.path.join(ROOT, "tests", "golden", "test_flame1.flam3")).read()
    b = resize(f1, wl["w"], wl["h"], scale=wl["scale"], quality=q, oversample=wl["oversample"],
               estimator_radius=9, temporal_samples=wl["temporal_samples"])
    return "<flames>" + a + b.replace('time="0"', 'time="100"', 1) + "</flames>"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks
    line). NVML is polled from a thread every few milliseconds (the timed region can be
    shorter than one nvidia-smi polling period); nvidia-smi -lms is the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu_index = [], None, gpu_index
        self.nvml, self.handle, self.thread, self.stop_flag = None, None, None, False
        self.sm, self.mx, self.reasons = [], [], set()

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu_index
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu_index])
                except Exception:
                    idx = self.gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = {n.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 n.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 n.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 n.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_ncu_traffic():
    """dram bytes per launch of the chaos-game kernel from the committed ncu --set full
    summary of this round (profiles/iterate_traffic.json says which capture), or None."""
    p = os.path.join(ROOT, "profiles", "iterate_traffic.json")
    try:
        return json.load(open(p)).get("dram_bytes_per_launch")
    except Exception:
        return None


def _ref_time(ref, cores, q, wl=None):
    wl = wl or WL
    g = ref.parse(workload_xml(q, wl))
    t = time.perf_counter()
    _, st, _ = ref.render(g, nthreads=cores, **{k: v for k, v in wl["render_kw"].items() if k != "nthreads"})
    return time.perf_counter() - t, st["num_iters"]


def _calibrate(ref, cores):
    """Reference frame time model t(q) = fixed + per_q * q at the workload's size, from
    short renders (the post-pass cost does not scale with quality); only used to SIZE the
    sample, never reported as a measurement."""
    q0, q1 = max(1.0, WL["quality"] / 200.0), max(4.0, WL["quality"] / 50.0)
    t0, _ = _ref_time(ref, cores, q0)
    t1, _ = _ref_time(ref, cores, q1)
    per_q = max(1e-6, (t1 - t0) / (q1 - q0))
    fixed = max(0.0, t0 - q0 * per_q)
    return per_q, fixed


def _quality_for(seconds, per_q, fixed):
    return float(max(1.0, min(WL["quality"], round((seconds - fixed) / per_q))))


def reference_arm(args, rank):
    """The UNMODIFIED reference (oracle/_ref) on all host cores, on this arm's own config:
    every step is one flam3_render of the same genome at the same size; when the full
    quality would not let warm-up + steps end within a few minutes the step is a bounded
    sample of the frame (reduced quality, said so in cpu_baseline.sample; the fixed post-pass
    cost then weighs ~5 % more per iteration than at full quality)."""
    if rank != 0:
        return
    import flam3_b200 as fb

    ref = fb.Flam3Lib(os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so"))
    cores = os.cpu_count() or 1
    per_q, fixed = _calibrate(ref, cores)
    budget = 170.0 / max(1, args.steps + args.warmup)
    q = _quality_for(budget, per_q, fixed)
    g = ref.parse(workload_xml(q))
    kw = {k: v for k, v in WL["render_kw"].items() if k != "nthreads"}
    for _ in range(args.warmup):
        ref.render(g, nthreads=cores, **kw)
    t = time.perf_counter()
    iters = 0.0
    for _ in range(args.steps):
        _, st, _ = ref.render(g, nthreads=cores, **kw)
        iters += st["num_iters"]
    dt = time.perf_counter() - t
    v = iters / dt
    full = q >= WL["quality"]
    sample = ("%s: same genome and size at quality %g (%.3g iters/step, %.1f s/step), unmodified reference "
              "flam3_render, nthreads=%d" % ("whole frame" if full else "bounded sample of the frame", q,
                                              iters / max(1, args.steps), dt / max(1, args.steps), cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (genome fixtures tests/golden/test_flame*.flam3)",
        "config": {"workload": WL["name"]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(seconds=30.0):
    """The reference on the host cores: the WHOLE frame when that fits ~30 s (C2 does: 10-20 s
    on 16-32 cores), else a bounded sample at reduced quality."""
    import flam3_b200 as fb

    ref = fb.Flam3Lib(os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so"))
    cores = os.cpu_count() or 1
    per_q, fixed = _calibrate(ref, cores)
    q = _quality_for(seconds, per_q, fixed)
    dt, iters = _ref_time(ref, cores, q)
    full = q >= WL["quality"]
    out = {"value": iters / dt, "unit": UNIT, "cores": cores, "kind": "reference",
           "sample": "%s: same genome and size at quality %g (%.3g iters, %.1f s), unmodified reference "
                     "flam3_render, nthreads=%d" % ("whole frame" if full else "bounded sample of the frame",
                                                    q, iters, dt, cores)}
    if full:
        out["frame_time_s"] = dt
    return out


def main():
    global WL
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    WL = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return 0

    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    import numpy as np
    import torch

    import flam3_b200 as fb

    os.environ["F3B_DEVICE"] = str(local_rank)
    os.environ.pop("F3B_DEVICES", None)
    if world > 1 and "F3B_HOST_THREADS" not in os.environ:
        # N processes share the host: each packs its temporal samples on its share of the cores
        os.environ["F3B_HOST_THREADS"] = str(max(1, min(16, (os.cpu_count() or 1) // world)))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    core = fb.load_core()
    gpu = fb.load_dropin()
    genomes = gpu.parse(workload_xml())
    kw = dict(WL["render_kw"])
    geo = gpu.buffer_geometry(genomes, **{k: v for k, v in kw.items() if k in ("time", "ngenomes")})
    nb = geo["width"] * geo["height"]
    frames = list(range(100)) if args.workload == "c4" else [None]
    animate = args.workload == "c4"

    def render(out, step=0):
        k = dict(kw)
        if animate:  # frames dealt round-robin to the ranks: no communication at all
            k["time"] = float(frames[(step * world + rank) % len(frames)])
        _, st, _ = gpu.render(genomes, out=out, **k)
        return st

    ctx = gpu.render_context(core)
    ctx.width, ctx.height = geo["width"], geo["height"]
    out = np.zeros((WL["h"], WL["w"], 3), dtype=np.uint8)
    gpu.set_shard(0, 1)
    if world > 1 and not animate:
        from flam3_b200.multigpu import setup_peer_strips

        setup_peer_strips(gpu, core, rank, world, genomes, dist=dist, **kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # ---- device-timed leg: images stay on the GPU ------------------------------------
    gpu.set_device_only(True)
    for i in range(args.warmup):
        st = render(out, i)
        assert st["rv"] == 0, st
    iters_per_step = st["num_iters"]
    barrier()
    ctx.stats_reset()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()  # starting the sampler takes rank 0 a while: everybody enters the timed region together
    ctx.timer_start()
    for i in range(args.steps):
        render(out, i)
    dev_ms = ctx.timer_stop()
    barrier()
    it_ms, it_n = ctx.stage_stats(0)
    de_ms, de_n = ctx.stage_stats(1)
    fi_ms, fi_n = ctx.stage_stats(2)
    ex_ms, ex_n = ctx.stage_stats(3)
    lost, saturated = ctx.exactness()
    clk = clocks.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())

    # ---- end-to-end leg: flam3_render with a host output buffer ------------------------
    gpu.set_device_only(False)
    for i in range(2):
        render(out, i)
    barrier()
    ctx.stats_reset()
    t0 = time.perf_counter()
    for i in range(args.steps):
        st = render(out, i)
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    h2d, d2h = ctx.xfer_stats()
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # ---- N > 1: the frame rendered together equals the frame rendered by one GPU --------
    identical = None
    if world > 1 and not animate:
        together = out.copy()
        barrier()
        if rank == 0:
            gpu.set_peers(None, None)
            gpu.set_shard(0, 1)
            alone = np.zeros_like(out)
            gpu.render(genomes, out=alone, **kw)
            identical = bool(np.array_equal(alone, together))
        barrier()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (chaos game + histogram) ----------------------
    # algorithmic traffic: 20 B read-modify-write per in-bounds hit; hits
    # counted from this rank's own buckets (count channel / 255).
    hist = ctx.hist_download_u32()
    hits_rank = float(hist[..., 4].astype(np.float64).sum() / 255.0)
    if world > 1 and identical is not None:
        hits_rank = hits_rank / world  # rank 0 just rendered the whole frame alone
    peak, peak_src = load_peaks()
    launches_iter = max(1, args.steps)
    iter_s_per_launch = (it_ms / 1e3) / launches_iter
    alg_bytes = 20.0 * hits_rank
    achieved = alg_bytes / iter_s_per_launch / 1e9
    roofline = {
        "kernel": "f3b_iterate_kernel (chaos game + red.global.add.v4.f32 into float staging planes + sweeper "
                  "warps; time includes f3b_hist_merge_kernel into the uint32 buckets)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": load_ncu_traffic(),
        "peak_source": peak_src,
        "note": ("the contract's HBM-equivalent figure: 20 B algorithmic read-modify-write per in-bounds hit "
                 "over the measured copy bandwidth. The kernel is an L2-reduction scatter (the staging planes "
                 "are L2-resident, DRAM sees a few % of these bytes); what bounds it is l2_red below."),
        "l2_red": {"bound": "L2 vector-reduction rate + L1TEX wavefronts (see DESIGN.md section 5)",
                   "achieved_hits_per_s": hits_rank / iter_s_per_launch,
                   "ceiling_hits_per_s": 1.95e11,
                   "ceiling_source": "tools/micro/red_peak.cu: uniform-random red.global.add.v4.f32 into a 34 MB "
                                     "plane, profiles/r01_red_peak_microbench.txt",
                   "frac": hits_rank / iter_s_per_launch / 1.95e11},
        "avg_launch_ms": 1e3 * iter_s_per_launch,
        "iters_per_s_kernel_only": iters_per_step * (world if animate else 1) / iter_s_per_launch,
        "in_bounds_fraction": hits_rank * (1 if animate else world) / iters_per_step,
        "share_of_step": it_ms / dev_ms,
    }
    # post-pass stages: 36 B/bucket for log scale + DE, 16 B/bucket + the
    # image for filter + tone map; with N ranks each one handles a 1/N row strip.
    post = []
    steps_timed = float(args.steps)
    strip = 1 if animate else world
    if de_n:
        b = nb * 36.0 / strip
        t_stage = de_ms / 1e3 / steps_timed
        post.append({"stage": "log scale + density estimation (f3b_de_classify_kernel + "
                              "f3b_de_gather_kernel)", "bound": "hbm",
                     "achieved": b / t_stage / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": b / t_stage / 1e9 / peak, "stage_ms": 1e3 * t_stage,
                     "launches_per_step": de_n / steps_timed})
    if fi_n:
        b = (nb * 16.0 + WL["w"] * WL["h"] * 3.0) / strip
        t_stage = fi_ms / 1e3 / steps_timed
        post.append({"stage": "spatial filter + tone map (f3b_final_kernel)", "bound": "hbm",
                     "achieved": b / t_stage / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": b / t_stage / 1e9 / peak, "stage_ms": 1e3 * t_stage,
                     "launches_per_step": fi_n / steps_timed})
    if ex_n:
        # rows of the strip (+ halo) fetched from world-1 peers, 16 B per bucket (+4 with a count plane)
        t_stage = ex_ms / 1e3 / steps_timed
        post.append({"stage": "strip reduction over peer memory (f3b_hist_reduce_kernel, NVLink loads)",
                     "stage_ms": 1e3 * t_stage, "launches_per_step": ex_n / steps_timed,
                     "nvlink_peak_GBps": 770.0})

    total_iters = iters_per_step * args.steps * (world if animate else 1)
    value = total_iters / (dev_ms / 1e3)
    staging_mb = 16.0 * nb / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if animate else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (genome fixtures tests/golden/test_flame*.flam3, random walk seeds)",
        "config": {"workload": WL["name"]},
        "workload_detail": {
            "iters_per_step": iters_per_step, "buckets": [geo["width"], geo["height"]],
            "l2_policy": "every step rewrites its whole working set (float staging copies of %.0f MB each, "
                         "uint32 buckets %.0f MB, accumulator %.0f MB, DE index plane): larger than the "
                         "126 MB L2 at this size; scatter addresses are random, nothing is reused across "
                         "steps" % (staging_mb, staging_mb, staging_mb),
            "parallelism": ("%d frames in flight, one per GPU, no communication" % world) if animate else
                           (("chains sharded over %d GPUs (one process each); fused strip reduction over peer "
                             "memory (NVLink loads, CUDA IPC), strip post-pass, image rows written into rank "
                             "0's image; barriers: stream-ordered 4-byte NCCL all-reduces" % world)
                            if world > 1 else "single GPU"),
        },
        "frame_time_s": dev_ms / 1e3 / args.steps,
        "e2e": {"value": total_iters / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "flam3_render(flam3_frame*, void *out, field, nchan, transp, stat_struct*)"},
        "gpu_launches": int(it_n + de_n + fi_n + ex_n),
        "exact_buckets": {"staged_values_at_2^24": int(lost), "saturated_uint32_sums": int(saturated)},
        "roofline": roofline, "roofline_post": post, "clocks": clk,
    }
    if animate:
        line["frames_per_s"] = args.steps * world / (dev_ms / 1e3)
    if identical is not None:
        line["image_identical_to_single_gpu"] = identical
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_leg()
        except Exception as e:  # the checker is missing: report, do not fake
            line["cpu_baseline"] = {"unavailable": str(e)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
