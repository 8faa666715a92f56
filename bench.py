#!/usr/bin/env python
"""bench.py — headline benchmark of the flam3 render hot path on B200.

Metric (BASELINE.json): chaos-game iterations/sec (and the 1080p q2000 frame time it
implies). A "step" is one full render of the workload frame: chaos game + histogram,
log-scale + density estimation, spatial filter + tone map.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload at every N: config C2 of SURVEY.md §8(d) — test.flam3's first flame at 1920x1080,
quality 2000, oversample 1, estimator_radius 9 (4.147e9 stored iterations per frame). With
N > 1 the frame's chains are split over the ranks (strong scaling), the private
histograms are summed with one NCCL all-reduce, every rank post-processes a strip of image
rows and an NCCL all-gather assembles the image (flam3_b200/multigpu.py).

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

W, H, QUALITY, SCALE = 1920, 1080, 2000.0, 540.0
WORKLOAD = "C2: test.flam3 flame 0 @1920x1080 q2000 os1 estimator_radius 9 (flam3_render, bits=33)"
METRIC = "chaos_game_iterations_per_sec"
UNIT = "iters/s"


def workload_xml(quality=QUALITY):
    from flam3_b200.genome_xml import resize

    xml = open(os.path.join(ROOT, "tests", "golden", "test_flame0.flam3")).read()
    # flam3-render forces ntemporal_samples = 1 for stills (flam3-render.c:235)
    return resize(xml, W, H, scale=SCALE, quality=quality, oversample=1, estimator_radius=9,
                  temporal_samples=1)


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
    summary (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "iterate_traffic.json")
    try:
        return json.load(open(p)).get("dram_bytes_per_launch")
    except Exception:
        return None


def _calibrate(ref, cores):
    """Reference frame time model t(q) = fixed + per_q * q at the workload's size, from two
    short renders (the post-pass cost does not scale with quality)."""
    ts = []
    for q in (10.0, 40.0):
        g = ref.parse(workload_xml(q))
        t = time.perf_counter()
        ref.render(g, nthreads=cores)
        ts.append(time.perf_counter() - t)
    per_q = max(1e-5, (ts[1] - ts[0]) / 30.0)
    fixed = max(0.0, ts[0] - 10.0 * per_q)
    # refine the slope on a ~3 s render (short ones overrate the per-iteration speed)
    q3 = _quality_for(3.0, per_q, fixed)
    g = ref.parse(workload_xml(q3))
    t = time.perf_counter()
    ref.render(g, nthreads=cores)
    t3 = time.perf_counter() - t
    if q3 > 40.0:
        per_q = max(per_q, (t3 - fixed) / q3)
    return per_q, fixed


def _quality_for(seconds, per_q, fixed):
    return float(max(10.0, min(QUALITY, round((seconds - fixed) / per_q))))


def reference_arm(args, rank):
    """Unmodified reference on the host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    import flam3_b200 as fb

    ref_path = os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so")
    ref = fb.Flam3Lib(ref_path)
    cores = os.cpu_count() or 1
    per_q, fixed = _calibrate(ref, cores)
    # every step ~10 s, capped so that warm-up + steps stay within a few minutes
    budget = min(10.0, 170.0 / max(1, args.steps + args.warmup))
    q = _quality_for(budget, per_q, fixed)
    g = ref.parse(workload_xml(q))
    for _ in range(args.warmup):
        ref.render(g, nthreads=cores)
    t = time.perf_counter()
    iters = 0.0
    for _ in range(args.steps):
        _, st, _ = ref.render(g, nthreads=cores)
        iters += st["num_iters"]
    dt = time.perf_counter() - t
    v = iters / dt
    sample = "same genome and size at quality %g (%.3g iters/step, %.1f s/step), flam3_render CPU path, nthreads=%d" % (
        q, iters / max(1, args.steps), dt / max(1, args.steps), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (genome fixture tests/golden/test_flame0.flam3)",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frame_time_s_extrapolated_q2000": fixed + per_q * QUALITY,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(seconds=15.0):
    import flam3_b200 as fb

    ref = fb.Flam3Lib(os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so"))
    cores = os.cpu_count() or 1
    per_q, fixed = _calibrate(ref, cores)
    q = _quality_for(seconds, per_q, fixed)
    g = ref.parse(workload_xml(q))
    t = time.perf_counter()
    _, st, _ = ref.render(g, nthreads=cores)
    dt = time.perf_counter() - t
    return {"value": st["num_iters"] / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "same genome and size at quality %g (%.3g iters, %.1f s), unmodified "
                      "reference flam3_render, nthreads=%d" % (q, st["num_iters"], dt, cores),
            "frame_time_s_extrapolated_q2000": fixed + per_q * QUALITY}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

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
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    core = fb.load_core()
    gpu = fb.load_dropin()
    genomes = gpu.parse(workload_xml())
    geo = gpu.buffer_geometry(genomes)
    nb = geo["width"] * geo["height"]

    gpu.set_shard(rank, world)
    if world > 1:
        from flam3_b200.multigpu import make_allreduce_exchange, make_strip_gather

        ordered = not os.environ.get("F3B_HOST_SYNC_COLLECTIVES")
        gpu.set_stream_ordered(ordered)
        gpu.set_exchange(make_allreduce_exchange(core, nb, stream_ordered=ordered))
        if not os.environ.get("F3B_NO_STRIPS"):
            gpu.set_strip_gather(make_strip_gather(rank, world, core=core, stream_ordered=ordered))

    ctx = gpu.render_context(core)
    ctx.width, ctx.height = geo["width"], geo["height"]
    out = np.zeros((H, W, 3), dtype=np.uint8)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # ---- device-timed leg: images stay on the GPU ------------------------------------
    gpu.set_device_only(True)
    for _ in range(args.warmup):
        _, st, _ = gpu.render(genomes, out=out)
        assert st["rv"] == 0, st
    iters_per_step = st["num_iters"]
    barrier()
    ctx.stats_reset()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ctx.timer_start()
    for _ in range(args.steps):
        gpu.render(genomes, out=out)
    dev_ms = ctx.timer_stop()
    barrier()
    it_ms, it_n = ctx.stage_stats(0)
    de_ms, de_n = ctx.stage_stats(1)
    fi_ms, fi_n = ctx.stage_stats(2)
    clk = clocks.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())

    # ---- end-to-end leg: flam3_render with a host output buffer ------------------------
    gpu.set_device_only(False)
    for _ in range(2):
        gpu.render(genomes, out=out)
    barrier()
    ctx.stats_reset()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, st, _ = gpu.render(genomes, out=out)
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    h2d, d2h = ctx.xfer_stats()
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (chaos game + histogram) ----------------------
    # algorithmic traffic: 20 B read-modify-write per in-bounds hit; hits
    # counted from the histogram itself (alpha plane / 255 when it doubles as the count).
    hist = ctx.hist_download_u32()
    hits_total = float(hist[..., 4].astype(np.float64).sum() / 255.0)
    peak, peak_src = load_peaks()
    # one chaos-game launch per step at this workload (+ the 40 us merge of the histogram
    # copies, which the stage time includes): per-launch figures are per step
    launches_iter = max(1, args.steps)
    iter_s_per_launch = (it_ms / 1e3) / launches_iter
    hits_per_launch = hits_total / 1.0 / world  # per rank and frame == per launch
    alg_bytes = 20.0 * hits_per_launch
    achieved = alg_bytes / iter_s_per_launch / 1e9
    roofline = {
        "kernel": "f3b_iterate_kernel<COUNT=false, SIMPLE=true> (chaos game + red.global.add.v4.f32 "
                  "histogram; time includes f3b_hist_merge_kernel)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": load_ncu_traffic(),
        "peak_source": peak_src,
        "note": ("L2-reduction / L1TEX-bound scatter, not an HBM stream: 20 B algorithmic RMW per "
                 "in-bounds hit; the histogram copies are L2-resident. See DESIGN.md section 5."),
        "avg_launch_ms": 1e3 * iter_s_per_launch,
        "iters_per_s_kernel_only": iters_per_step / world / iter_s_per_launch * world,
        "in_bounds_fraction": hits_total / iters_per_step,
        "share_of_step": it_ms / dev_ms,
    }
    # post-pass stages: 36 B/bucket for log scale + DE, 16 B/bucket + the
    # image for filter + tone map; with N ranks each one handles a 1/N row strip. These are
    # per-stage figures (DE = classify + gather launches); at this size both stages are
    # bound by their arithmetic (DE taps, fp64 pow), not by HBM: see DESIGN.md section 5.
    post = []
    steps_timed = float(args.steps)
    if de_n:
        b = nb * 36.0 / world
        t_stage = de_ms / 1e3 / steps_timed
        post.append({"stage": "log scale + density estimation (f3b_de_classify_kernel + "
                              "f3b_de_gather_kernel)", "bound": "hbm",
                     "achieved": b / t_stage / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": b / t_stage / 1e9 / peak, "stage_ms": 1e3 * t_stage,
                     "launches_per_step": de_n / steps_timed})
    if fi_n:
        b = (nb * 16.0 + W * H * 3.0) / world
        t_stage = fi_ms / 1e3 / steps_timed
        post.append({"stage": "spatial filter + tone map (f3b_final_kernel)", "bound": "hbm",
                     "achieved": b / t_stage / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": b / t_stage / 1e9 / peak, "stage_ms": 1e3 * t_stage,
                     "launches_per_step": fi_n / steps_timed})

    total_iters = iters_per_step * args.steps
    value = total_iters / (dev_ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (genome fixture tests/golden/test_flame0.flam3, random walk seeds)",
        "config": {"workload": WORKLOAD, "iters_per_step": iters_per_step,
                   "buckets": [geo["width"], geo["height"]],
                   "l2_policy": "every step clears and rewrites its whole working set: 3 histogram "
                                "copies (102 MB) + accumulator (34 MB) + DE index plane (4 MB) = 140 MB, "
                                "larger than the 126 MB L2; scatter addresses are random, nothing is "
                                "reused across steps",
                   "parallelism": ("chains sharded over %d GPUs, NCCL all-reduce of the histograms, "
                                   "row-strip post-pass + NCCL all-gather of the image strips" % world)
                   if world > 1 else "single GPU"},
        "frame_time_s": dev_ms / 1e3 / args.steps,
        "e2e": {"value": total_iters / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "flam3_render(flam3_frame*, void *out, field, nchan, transp, stat_struct*)"},
        "gpu_launches": int(it_n + de_n + fi_n),
        "roofline": roofline, "roofline_post": post, "clocks": clk,
    }
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
