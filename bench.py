#!/usr/bin/env python
"""bench.py -- exact-OIT frames/s of the B200 rasteriser on BASELINE.json's synthetic configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config I] [--impl ours|reference] [--mode split|views]

A "step" is one frame: quad setup -> binning -> raster low/high -> image.  The workload is the configuration
BASELINE.json's target is quoted on: configs[3], the 10M-triangle textured architecture scene at 3840x2160.
  N = 1   one GPU renders the whole frame;
  N > 1   `--mode split` (the default): the bin rows of that SAME frame are split over the ranks as
          cost-balanced row-major bin ranges, every rank stores its strip straight into rank 0's image over
          NVLink ("scaling": "strong").  The line also carries "views": the 64-view orbit of the 1M-triangle
          scene (configs[4]) with the views sharded over the ranks, no data-path collective.

`value` is frames/s with geometry resident in HBM, timed with CUDA events on the stream the kernels are launched
on.  The frame's own per-frame inputs (352-byte LucidConfig as kernel arguments, 36 bytes per instance) are
submitted inside the bracket, but the instance copy runs on the library's upload stream at submission time and
the host runs up to three frames ahead, so in steady state that copy overlaps the previous frame.  `e2e` is the
same frame through lucid_render() with HOST buffers in and the RGBA8 image + LucidInfo copied back into pinned
host memory inside the timed region (wall clock).  `sustained` is at least a second of back-to-back frames by
wall clock.  `--impl reference` times the CPU restatement of the reference shaders (oracle/) with all host
threads: the reference's own Vulkan path cannot run on this image (no ICD, no shaderc).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    0: "config0: 100k-triangle quad soup, 50% alpha=0.5, 1280x720",
    1: "config1: 1M-triangle meshlet scene (489 patches of 32x32 quads), mixed opaque/transparent, 1920x1080",
    2: "config2: hairball-like dense overlap (median depth complexity 72), 5M triangles, 3840x2160",
    3: "config3: 10M-triangle architecture scene, textured atlas shading, 3840x2160",
    4: "config4: 64-view orbit of the 1M-triangle scene, 1920x1080",
}
DEFAULT_CONFIG = 3  # BASELINE.json: "4K exact-OIT frames/sec on a 10M-triangle synthetic scene"
KERNEL_SOURCES = ["lucid_b200/csrc/common.cuh", "lucid_b200/csrc/setup.cu", "lucid_b200/csrc/binning.cu",
                  "lucid_b200/csrc/raster_common.cuh", "lucid_b200/csrc/raster_bins.cu", "lucid_b200/csrc/raster_sort.cu",
                  "lucid_b200/csrc/raster_shade.cu"]


def kernel_source_hash() -> str:
    """Digest of the kernel sources: profiles/dram_traffic.json records the one its ncu capture was taken at."""
    h = hashlib.sha256()
    for rel in KERNEL_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


_SAMPLER = r"""
import json, sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
sm, bits = [], 0
print("ready", flush=True)
import select
while True:
    if select.select([sys.stdin], [], [], 0.005)[0]:
        break
    try:
        sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
    except Exception:
        pass
print(json.dumps({"sm": sm, "bits": bits, "max": mx}), flush=True)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~5 ms while the timed region runs.  The poller
    is a separate PROCESS: as a thread of this one it took the GIL (and NVML's locks) in the middle of a frame's
    launches, which showed up as millisecond outliers in the multi-GPU step times."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER, str(self.phys)], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            if self.proc.stdout.readline().strip() != "ready":
                self.proc = None
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        try:
            out, _ = self.proc.communicate("stop\n", timeout=5.0)
            d = json.loads(out.strip().splitlines()[-1])
        except Exception:
            self.proc.kill()
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted(n for n, bit in self.REASONS.items() if d["bits"] & bit)
        return {"sm_mhz": float(np.median(d["sm"])) if d["sm"] else None, "sm_max_mhz": d["max"], "reasons": reasons,
                "samples": len(d["sm"])}


def make_scene(config: int, scale: float):
    from lucid_b200 import scenes
    return scenes.get_config(config, scale)


def view_camera(scene, view: int, num_views: int = 64):
    cam = dict(scene["camera"])
    if cam["kind"] == "orbit" and view:
        cam["rot_h"] = cam["rot_h"] + 2.0 * np.pi * view / num_views
    return cam


def algorithmic_bytes(stats: dict, scene, width, height):
    """SURVEY.md 8(d) byte formulas with this implementation's record sizes."""
    n_in = stats["input_quads"]
    n_vis = stats["visible_small"] + stats["visible_large"]
    n_bq, n_bt = stats["bin_quads"], stats["bin_tris"]
    t_bin = 2 * n_bq + n_bt
    attr = (16 if scene.get("colors") is not None else 0) + (16 if scene.get("normals") is not None else 0) + \
           (32 if scene.get("uvs") is not None else 0)
    n_px = (stats["low_bins"] + stats["high_bins"]) * 1024
    setup = 64 * n_in + (4 + 2 * 96 + attr) * n_vis
    count = 4 * n_vis + 32 * 2 * stats["visible_large"]
    dispatch = count + 4 * (n_bq + n_bt)
    raster = 4 * (n_bq + n_bt) + (96 + attr / 2) * t_bin + 4 * n_px
    return {"setup": setup, "bin_count": count, "bin_dispatch": dispatch, "raster": raster}


def depth_complexity(frag_counts: np.ndarray) -> dict:
    cov = frag_counts[frag_counts > 0]
    if cov.size == 0:
        return {"covered_frac": 0.0, "median": 0, "p99": 0, "max": 0}
    return {"covered_frac": round(float(cov.size) / frag_counts.size, 4), "median": float(np.median(cov)),
            "p99": float(np.percentile(cov, 99)), "max": int(cov.max())}


class Rig:
    """One renderer on this rank's GPU with a scene resident, plus what the timed loops share."""

    def __init__(self, args, config, torch, dist, rank, world, local_rank, stream, scene=None):
        from lucid_b200 import api
        self.api, self.torch, self.dist = api, torch, dist
        self.rank, self.world = rank, world
        self.scene = scene if scene is not None else make_scene(config, args.scale)
        self.width, self.height = self.scene["width"], self.scene["height"]
        self.stream = stream
        self._configs = {}
        self.r = api.LucidRenderer(self.width, self.height, 0, args.mvq, device=local_rank, stream=stream.cuda_stream)
        self.r.set_scene(self.scene)
        self.inst, self.cols, self.rects = api.build_instances(self.scene["draw_calls"], self.scene["materials"])
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        self.host_imgs = [torch.empty((self.height, self.width), dtype=torch.int32).pin_memory() for _ in range(2)]

    def config_for(self, view):
        api = self.api
        if view not in self._configs:  # a handful of cameras, asked for every frame
            cam = api.make_camera(view_camera(self.scene, view), self.width, self.height)
            self._configs[view] = api.make_config(cam, len(self.inst), self.scene["background"])
        return self._configs[view]

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def close(self):
        self.r.close()


def timed_steps(rig, args, render, after_frame, steps):
    """W warm-up steps, then exactly `steps` steps between a barrier + synchronize on both sides; every step is
    bracketed by CUDA events on the launching stream, L2 flushed (untimed) between steps.  Returns the per-step
    times of this rank (ms) and the wall time of the loop."""
    torch = rig.torch
    plain = rig.api.RENDER_ASYNC | rig.api.RENDER_SKIP_INFO | rig.api.RENDER_NO_STAGE_TIMES | rig.api.RENDER_CULL_INSTANCES
    rig.barrier()  # every rank enters the loop together (the device-side frame flags give up after a few seconds)
    for w in range(args.warmup):  # the same sequence as a timed step
        render(w, plain)
        after_frame()
    rig.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t_wall = time.time()
    for k in range(steps):
        rig.flush.fill_(k & 0xFF)  # L2 flush between timed iterations (untimed)
        starts[k].record(rig.stream)
        render(args.warmup + k, plain)
        after_frame()
        stops[k].record(rig.stream)
    rig.barrier()
    wall = time.time() - t_wall
    return np.array([s.elapsed_time(e) for s, e in zip(starts, stops)], np.float64), wall


def max_over_ranks(rig, value: float) -> float:
    if rig.dist is None:
        return value
    t = rig.torch.tensor([value], device="cuda", dtype=rig.torch.float64)
    rig.dist.all_reduce(t, op=rig.dist.ReduceOp.MAX)
    return float(t.item())


def sustained_run(rig, render, after_frame, seconds: float, frames_per_step: int):
    """Back-to-back frames for at least `seconds` of wall clock, no flush in between (the scene's records are
    several times the L2), device-timed by one event pair around the whole run."""
    torch = rig.torch
    plain = rig.api.RENDER_ASYNC | rig.api.RENDER_SKIP_INFO | rig.api.RENDER_NO_STAGE_TIMES | rig.api.RENDER_CULL_INSTANCES
    rig.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(rig.stream)
    n = 0
    while True:
        for _ in range(16):
            render(n, plain)
            after_frame()
            n += 1
        done = time.perf_counter() - t0 >= seconds  # the host is never more than three frames ahead of the GPU
        if rig.dist is not None:  # every rank runs the same number of frames: rank 0's clock decides
            stop = rig.torch.tensor([1.0 if done else 0.0], device="cuda")
            rig.dist.broadcast(stop, src=0)
            done = float(stop.item()) > 0
        if done:
            break
    e1.record(rig.stream)
    rig.barrier()
    wall = time.perf_counter() - t0
    dev_ms = max_over_ranks(rig, e0.elapsed_time(e1))
    return {"frames": n * frames_per_step, "wall_s": round(wall, 3), "device_s": round(dev_ms / 1e3, 3),
            "value": round(n * frames_per_step / (dev_ms / 1e3), 3), "unit": "frames/s",
            "note": "back-to-back frames, no L2 flush (per-frame records exceed the L2), one CUDA event pair"}


def e2e_run(rig, frame, steps, frames_per_step, renderers=None):
    renderers = renderers or [rig.r]
    rig.barrier()
    for w in range(max(3, len(renderers))):
        frame(w)
    for r in renderers:
        r.wait()
    rig.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        frame(k)
    for r in renderers:
        r.wait()
    rig.barrier()
    return frames_per_step * steps / max_over_ranks(rig, time.perf_counter() - t0)


def run_views(args, torch, dist, rank, world, local_rank, stream, config):
    """Multi-view batch: view v of the orbit goes to rank v % world; no data-path communication."""
    rig = Rig(args, config, torch, dist, rank, world, local_rank, stream)
    api, r = rig.api, rig.r

    def render(step, flags):
        r.render(rig.config_for((step * world + rank) % 64), rig.inst, rig.cols, rig.rects, flags=flags)

    step_ms, _ = timed_steps(rig, args, render, lambda: None, args.steps)
    total_ms = max_over_ranks(rig, float(step_ms.sum()))
    e2e_steps = max(args.steps, 3)

    def e2e_frame(k):
        r.render(rig.config_for((k * world + rank) % 64), rig.inst, rig.cols, rig.rects,
                 out=rig.host_imgs[k & 1].data_ptr(), flags=api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES)

    e2e = e2e_run(rig, e2e_frame, e2e_steps, world)
    out = {"workload": WORKLOADS[4], "value": round(world * 1000.0 * args.steps / total_ms, 3), "unit": "frames/s",
           "ms_per_step": round(total_ms / args.steps, 4), "scaling": "weak",
           "e2e": {"value": round(e2e, 3), "unit": "frames/s", "h2d_bytes_per_step": int(len(rig.inst) * 36 + 352),
                   "d2h_bytes_per_step": int(rig.width * rig.height * 4 + (1152 + 10 * r.bin_count) * 4)}}
    rig.close()
    return out


def run_ours(args):
    import torch
    from lucid_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL announces its version on stdout at the first communicator: keep stdout to the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    # a stream of our own, made torch's current stream: the library launches every kernel on the stream it
    # is handed (handle 0 would make it create a private one, invisible to torch.cuda.Event), so the
    # flush, the timing events and the kernels are all ordered on this one stream
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    mode = args.mode or ("split" if world > 1 else "single")
    if world == 1:
        mode = "single"
    if mode == "views":  # the secondary workload on its own (configs[4]); the line is that run's
        v = run_views(args, torch, dist, rank, world, local_rank, stream, 1 if args.config == DEFAULT_CONFIG else args.config)
        if rank == 0:
            line = {"metric": "exact_oit_frames_per_sec", "value": v["value"], "unit": "frames/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": v["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": v["workload"], "parallelism": "views sharded x%d" % world,
                               "l2": "256 MiB device memset between timed frames (untimed)"},
                    "e2e": v["e2e"], "gpu_launches": int(9 * args.steps)}
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    split = mode == "split"
    rig = Rig(args, args.config, torch, dist, rank, world, local_rank, stream)
    r, scene, width, height = rig.r, rig.scene, rig.width, rig.height
    inst, cols, rects = rig.inst, rig.cols, rig.rects
    nby = (height + 31) // 32
    rows, split_kind = None, None
    if split:
        # ownership from the measured raster cost of a full calibration frame (what an application takes
        # from the previous frame); rank 0's measurement is used by every rank.  Default: contiguous
        # row-major bin ranges of equal cost (a heavy bin row may be shared by two ranks); --split-rows
        # keeps whole bin rows, --equal-rows equal row counts.
        from lucid_b200 import multigpu
        for _ in range(2):
            r.render(rig.config_for(0), inst, cols, rects)
        cost = torch.from_numpy(r.read_bin_costs().astype(np.float64)).cuda()
        dist.broadcast(cost, src=0)
        cost = cost.cpu().numpy()
        if args.equal_rows or args.split_rows:
            weights = None if args.equal_rows else cost.reshape(nby, -1).sum(axis=1)
            rows = multigpu.split_bin_rows(nby, world, weights)[rank]
            r.set_bin_rows(*rows)
            split_kind = "equal rows" if args.equal_rows else "whole rows balanced on measured cost"
        else:
            ranges = multigpu.split_bins(r.bin_count, world, cost)
            r.set_bin_range(*ranges[rank])
            # feedback, as an application would apply it from frame to frame: the cost of a rank's bins is
            # rescaled by the time that rank's whole frame actually took, and the ranges are cut again (seven
            # rounds bring the slowest rank within 2 % of the mean on the 10M-triangle scene, tools/split_probe.py)
            for _ in range(args.balance_iters):
                for k in range(3):
                    r.render(rig.config_for(0), inst, cols, rects,
                             flags=api.RENDER_ASYNC | api.RENDER_SKIP_INFO | api.RENDER_CULL_INSTANCES)
                mine = float(np.median([r.stage_times(i)[7] for i in range(2)]))  # the rank's whole frame
                times = torch.zeros(world, device="cuda", dtype=torch.float64)
                times[rank] = mine
                dist.all_reduce(times)
                times = times.cpu().numpy()
                for q, (lo, hi) in enumerate(ranges):
                    cost[lo:hi] *= times[q] / max(float(cost[lo:hi].sum()), 1e-9)
                ranges = multigpu.split_bins(r.bin_count, world, cost)
                r.set_bin_range(*ranges[rank])
            rows = ranges[rank]
            split_kind = "row-major bin ranges balanced on measured cost, %d feedback steps" % args.balance_iters

    # composite target for the bin-row split: every rank stores into rank 0's image over NVLink, and signals the
    # frame's completion through a flag in rank 0's memory (lucid_signal / lucid_wait_flags); --completion allreduce
    # is the NCCL alternative.  A lane = one renderer handle on its own stream with its own gathered image and
    # flags on rank 0; lane 0 is the rig's handle, further lanes carry the frames in flight (below).
    use_flags = split and args.completion == "flags"
    frame_token = torch.zeros(1, device="cuda")
    RELEASED = api.LucidRenderer.SYNC_RELEASED

    class Lane:
        def __init__(self, lane_rig):
            self.rig, self.r, self.stream = lane_rig, lane_rig.r, lane_rig.stream
            self.peer_ptr, self.flags_ptr, self.frame_no = None, None, 0
            if split:
                handle = [(self.r.ipc_export_image(), self.r.ipc_export_sync()) if rank == 0 else None]
                dist.broadcast_object_list(handle, src=0)
                if rank != 0:
                    self.peer_ptr = self.r.ipc_open_image(handle[0][0])
                    self.flags_ptr = self.r.ipc_open_image(handle[0][1])
                else:
                    self.flags_ptr = self.r.sync_pointer()

        def render(self, step, flags):
            r = self.r
            if use_flags:
                self.frame_no += 1
                if rank != 0:  # the shared image may be stored into once the previous frame was released
                    r.set_frame_gate(self.flags_ptr, RELEASED, self.frame_no - 1)
            if self.peer_ptr is not None and args.composite == "stores":
                # the raster kernels store their pixels straight into rank 0's image
                r.render(rig.config_for(view_of(step)), inst, cols, rects, out_device_ptr=self.peer_ptr,
                         out_pitch=width * 4, flags=flags)
            else:
                r.render(rig.config_for(view_of(step)), inst, cols, rects, flags=flags)
                if self.peer_ptr is not None:  # own image first, then the owned bins as whole 128-byte rows
                    r.composite_to(self.peer_ptr, width * 4)

        def after_frame(self, consume=None):
            """The frame is complete when every rank's strip has landed in rank 0's image."""
            r = self.r
            if use_flags:
                if rank != 0:
                    r.signal(self.flags_ptr, rank, self.frame_no)
                else:
                    r.wait_flags(self.flags_ptr, 1, world - 1, self.frame_no)
                    if consume is not None:
                        consume()
                    r.signal(self.flags_ptr, RELEASED, self.frame_no)
            elif split:
                dist.all_reduce(frame_token)
                if consume is not None and rank == 0:
                    consume()

        def close(self):
            if self.peer_ptr is not None:
                self.r.ipc_close_image(self.peer_ptr)
                self.peer_ptr = None
            if rank != 0 and self.flags_ptr is not None:
                self.r.ipc_close_image(self.flags_ptr)
            self.flags_ptr = None

    def view_of(step):
        return 0 if split else step % 64

    lane0 = Lane(rig)
    render, after_frame = lane0.render, lane0.after_frame

    # NVML is polled by a separate process, for the GPU of the rank that prints the line
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms, wall = timed_steps(rig, args, render, after_frame, args.steps)
    kept = min(args.steps, 64)
    frame_ms = float(np.mean([r.stage_times(i)[7] for i in range(kept)]))  # library's own first->last event
    rank_frame_ms = None
    if dist is not None:  # every rank's own frame time (the split's balance; a step ends with the slowest)
        t = torch.zeros(world, device="cuda", dtype=torch.float64)
        t[rank] = frame_ms
        dist.all_reduce(t)
        rank_frame_ms = [round(float(x), 4) for x in t.cpu().numpy()]
    if split and args.trace_split:
        sys.stderr.write(f"[trace-split] rank {rank}: library frame {frame_ms:.4f} ms, step median "
                         f"{np.median(step_ms):.4f} ms; steps {[round(float(x), 3) for x in step_ms]}\n")
    total_ms = max_over_ranks(rig, float(step_ms.sum()))
    step_median = max_over_ranks(rig, float(np.median(step_ms)))

    sustained = sustained_run(rig, render, after_frame, args.sustained_seconds, 1)

    # Frames in flight: the kernels of a frame differ in what bounds them (setup: HBM; shading: instruction issue) and
    # each ends with a tail of few busy SMs -- long tails on a device that owns an eighth of the bins (one bin per
    # CTA, one list per warp).  F handles on F streams render alternate frames and the kernels of neighbouring frames
    # fill each other's gaps.  The same K steps, timed as one bracket; no L2 flush in between (a frame reads several
    # hundred MB of geometry and records, the L2 holds 126 MB).  Programmatic dependent launch is off here
    # (LUCID_RENDER_NO_DEPENDENT_LAUNCH): CTAs launched early wait on the SMs in the other handles' way.
    inflight = None
    if split:
        n_lanes = args.frames_in_flight if use_flags else 1
    else:
        n_lanes = args.frames_in_flight_single
    lanes = [lane0]
    if n_lanes > 1:
        for _ in range(n_lanes - 1):
            lane_stream = torch.cuda.Stream(device=local_rank)
            lr = Rig(args, args.config, torch, dist, rank, world, local_rank, lane_stream, scene=scene)
            if split and (args.equal_rows or args.split_rows):
                lr.r.set_bin_rows(*rows)
            elif split:
                lr.r.set_bin_range(*rows)
            lanes.append(Lane(lr))
        # several handles busy on one device: no programmatic dependent launch (include/lucid_b200.h)
        plain = (api.RENDER_ASYNC | api.RENDER_SKIP_INFO | api.RENDER_NO_STAGE_TIMES | api.RENDER_CULL_INSTANCES |
                 api.RENDER_NO_DEPENDENT_LAUNCH)

        def pipelined(steps):
            rig.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(rig.stream)
            for ln in lanes[1:]:
                ln.stream.wait_event(e0)
            for k in range(steps):
                ln = lanes[k % n_lanes]
                ln.render(k, plain)
                ln.after_frame()
            for ln in lanes[1:]:
                done = torch.cuda.Event()
                done.record(ln.stream)
                rig.stream.wait_event(done)
            e1.record(rig.stream)
            rig.barrier()
            return e0.elapsed_time(e1)

        pipelined(max(args.warmup, n_lanes) * n_lanes)
        # The ranges were balanced on frames rendered one at a time; with frames in flight the ranks gain differently
        # (a rank of many light bins overlaps better than one of few heavy bins), so the same feedback runs again on
        # each rank's own pipelined frame time: frames into the rank's own image, no hand-over, ranks uncoupled.
        if split and not (args.equal_rows or args.split_rows):
            local = plain
            cfg0 = rig.config_for(0)

            def standalone(frames):
                rig.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(rig.stream)
                for ln in lanes[1:]:
                    ln.stream.wait_event(e0)
                for k in range(frames):
                    lanes[k % n_lanes].r.render(cfg0, inst, cols, rects, flags=local)
                for ln in lanes[1:]:
                    done = torch.cuda.Event()
                    done.record(ln.stream)
                    rig.stream.wait_event(done)
                e1.record(rig.stream)
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / frames

            for _ in range(args.balance_iters_in_flight):
                standalone(2 * n_lanes)
                mine = standalone(6 * n_lanes)
                times = torch.zeros(world, device="cuda", dtype=torch.float64)
                times[rank] = mine
                dist.all_reduce(times)
                times = times.cpu().numpy()
                for q, (lo, hi) in enumerate(ranges):
                    cost[lo:hi] *= times[q] / max(float(cost[lo:hi].sum()), 1e-9)
                ranges = multigpu.split_bins(r.bin_count, world, cost)
                rows = ranges[rank]
                for ln in lanes:
                    ln.r.set_bin_range(*rows)
            pipelined(max(args.warmup, n_lanes) * n_lanes)
        pipe_ms = max_over_ranks(rig, pipelined(args.steps))
        t0 = time.perf_counter()
        n_sus, sus_ms = 0, 0.0
        while time.perf_counter() - t0 < args.sustained_seconds:  # rank 0's clock decides for everybody
            sus_ms += pipelined(16 * n_lanes)
            n_sus += 16 * n_lanes
            stop = torch.tensor([1.0 if time.perf_counter() - t0 >= args.sustained_seconds else 0.0], device="cuda")
            if dist is not None:
                dist.broadcast(stop, src=0)
            if float(stop.item()) > 0:
                break
        sus_ms = max_over_ranks(rig, sus_ms)
        inflight = {"frames_in_flight": n_lanes, "ms_per_step": pipe_ms / args.steps,
                    "value": 1000.0 * args.steps / pipe_ms,
                    "sustained": {"frames": n_sus, "device_s": round(sus_ms / 1e3, 3),
                                  "value": round(n_sus / (sus_ms / 1e3), 3), "unit": "frames/s"}}
    clocks = sampler.stop()

    # per-stage CUDA-event times: the same frames again, same L2 flush, this time with an event after every
    # stage (recorded on the stream the kernels are launched on; the library keeps its last 64 frames)
    rig.barrier()
    for k in range(kept):
        rig.flush.fill_(k & 0xFF)
        render(args.warmup + k, api.RENDER_ASYNC | api.RENDER_SKIP_INFO | api.RENDER_CULL_INSTANCES)
        after_frame()
    rig.barrier()
    stage = np.mean([r.stage_times(i).astype(np.float64) for i in range(kept)], axis=0)

    # end to end through the C ABI: host instance arrays in (the library stages them through pinned
    # memory), RGBA8 image + LucidInfo read back into pinned host memory every frame.  Frames are
    # submitted asynchronously: the copy-out of frame n overlaps the rendering of frame n+1.
    e2e_steps = max(args.steps, 3)

    def e2e_frame(k):
        if split:
            # every rank rasterises its bin ranges into rank 0's image, then rank 0 reads it back
            render(k, api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES | api.RENDER_CULL_INSTANCES)

            def read_back():
                torch.cuda.current_stream().synchronize()
                r.read_image_into(rig.host_imgs[k & 1].data_ptr())

            after_frame(read_back)
        else:
            r.render(rig.config_for(view_of(k)), inst, cols, rects, out=rig.host_imgs[k & 1].data_ptr(),
                     flags=api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES)

    e2e_value = e2e_run(rig, e2e_frame, e2e_steps, 1)
    e2e_extra = {}
    lane_renderers = [ln.r for ln in lanes]
    if not split and n_lanes > 1:
        # the same with the frames in flight: alternate frames on the handles, each with its own pair of host images
        e2e_extra["one_handle"] = {"value": round(e2e_value, 3), "unit": "frames/s"}
        e2e_flags = api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES | api.RENDER_NO_DEPENDENT_LAUNCH

        def e2e_frame_lanes(k):
            ln = lanes[k % n_lanes]
            ln.r.render(rig.config_for(view_of(k)), inst, cols, rects,
                        out=ln.rig.host_imgs[(k // n_lanes) & 1].data_ptr(), flags=e2e_flags)

        lanes_value = e2e_run(rig, e2e_frame_lanes, e2e_steps, 1, lane_renderers)
        if lanes_value > e2e_value:
            e2e_value = lanes_value
            e2e_extra["frames_in_flight"] = n_lanes
        else:  # the faster mode's figure is the line's, the other stays beside it
            e2e_extra["frames_in_flight"] = 1
            e2e_extra["with_%d_frames_in_flight" % n_lanes] = {"value": round(lanes_value, 3), "unit": "frames/s"}
    if split:
        # The frame gathered in HOST memory instead: one image shared by the processes (POSIX shared memory, pinned
        # in every process), every rank copies its own bins into it over its own PCIe link
        # (LUCID_RENDER_OWNED_BINS_ONLY) -- no device gathers the frame first, the 33 MB read-back of a 4K frame is
        # spread over the ranks' links.  Checked against the frame gathered on rank 0.
        from lucid_b200 import multigpu
        e2e_extra["gathered_on_rank0"] = {"value": round(e2e_value, 3), "unit": "frames/s",
                                          "d2h_bytes_per_step": int(width * height * 4)}
        render(0, api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES | api.RENDER_CULL_INSTANCES)
        after_frame()
        rig.barrier()
        ref_img = r.read_image() if rank == 0 else None
        name = "lucid_b200_e2e_%s" % os.environ.get("MASTER_PORT", "0")
        shared, problem = None, ""
        n_shared = 2 * n_lanes
        # a box that refuses the shared mapping or its page-locking (no /dev/shm, a memlock limit) must not cost the
        # line: every rank tries, the ranks agree, and without the shared image the rank-0 gather stays the e2e figure
        try:
            if rank == 0:
                shared = multigpu.SharedHostImages(name, width, height, n_shared, create=True)
        except Exception as e:  # noqa: BLE001
            problem = repr(e)
        rig.barrier()
        try:
            if rank != 0:
                shared = multigpu.SharedHostImages(name, width, height, n_shared, create=False)
            if shared is not None:
                shared.pin()
        except Exception as e:  # noqa: BLE001
            problem = repr(e)
        ok = torch.tensor([0.0 if (problem or shared is None) else 1.0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) > 0:
            host_flags = (api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES | api.RENDER_CULL_INSTANCES | api.RENDER_OWNED_BINS_ONLY |
                          (api.RENDER_NO_DEPENDENT_LAUNCH if n_lanes > 1 else 0))

            def e2e_frame_host(k):
                lanes[k % n_lanes].r.render(rig.config_for(view_of(k)), inst, cols, rects, out=shared.pointer(k % n_shared),
                                            flags=host_flags)

            e2e_value = e2e_run(rig, e2e_frame_host, e2e_steps, 1, lane_renderers)
            e2e_extra["frames_in_flight"] = n_lanes
            if rank == 0:
                e2e_extra["delivery"] = ("every rank copies its own bins into one host image shared by the processes "
                                         "(/dev/shm, page-locked), over its own PCIe link")
                e2e_extra["verified_against_gathered_frame"] = bool(
                    np.array_equal(shared.array[(e2e_steps - 1) % n_shared], ref_img))
        else:
            e2e_extra["delivery"] = "gathered on rank 0 and read back there (no shared host image on this box: %s)" % (problem or "another rank failed")
        rig.barrier()
        if shared is not None:
            shared.close()

    # counters of one frame for the roofline arithmetic, and the frame's depth complexity
    r.render(rig.config_for(0), inst, cols, rects, flags=api.RENDER_FRAG_COUNTS)
    info = r.read_info()
    stats = api.decode_stats(info, r.bin_count, width, height)
    depth = depth_complexity(r.read_frag_counts()) if not split else None

    tris_per_frame = 2 * stats["input_quads"]
    serial_ms_per_step = total_ms / args.steps
    frames_in_flight_tried = None
    if inflight is not None and not split and inflight["ms_per_step"] >= serial_ms_per_step:
        # one GPU, whole frames: frames in flight pay where a frame's kernels leave gaps to fill (the 10M-triangle
        # scene +5 %, the 1M-triangle scene +27 %) and not where one kernel after the other already fills the GPU
        # (hairball -4 %); an application would measure both once, and so does this run: the line is the faster mode's
        frames_in_flight_tried = {"frames_in_flight": inflight["frames_in_flight"],
                                  "value": round(inflight["value"], 3), "ms_per_step": round(inflight["ms_per_step"], 4)}
        inflight = None
    ms_per_step = serial_ms_per_step if inflight is None else inflight["ms_per_step"]
    value = 1000.0 / ms_per_step

    for ln in lanes[1:]:  # the extra handles of the frames in flight
        ln.close()
        ln.rig.close()
    views = None
    if split and not args.no_views:
        lane0.close()
        rig.close()
        views = run_views(args, torch, dist, rank, world, local_rank, stream, 1)

    if rank == 0:
        peak, peak_src = load_peaks()
        ab = algorithmic_bytes(stats, scene, width, height)
        # LOW and HIGH bins share the three raster kernels: block lists, block sort, shading
        stage_names = ["setup", "bin_count", "bin_scan", "bin_dispatch", "raster_lists", "raster_sort", "raster_shade"]
        stage_ms = {n: round(float(stage[i]), 4) for i, n in enumerate(stage_names)}
        stage_ms["frame_with_stage_events"] = round(float(stage[7]), 4)
        stage_ms["frame"] = round(frame_ms, 4)  # timed frames: first launch -> last kernel done, the library's own events
        # the library's events and ours are on one stream: our bracket can only be the wider one
        assert serial_ms_per_step >= 0.98 * frame_ms, (serial_ms_per_step, frame_ms)
        raster_ms = float(stage[4] + stage[5] + stage[6])
        fracs = {
            "setup": ab["setup"] / (stage[0] * 1e-3) / 1e9 / peak if stage[0] > 0 else None,
            "bin_count": ab["bin_count"] / (stage[1] * 1e-3) / 1e9 / peak if stage[1] > 0 else None,
            "bin_dispatch": ab["bin_dispatch"] / (stage[3] * 1e-3) / 1e9 / peak if stage[3] > 0 else None,
            "raster": ab["raster"] / (raster_ms * 1e-3) / 1e9 / peak if raster_ms > 0 else None,
        }
        dominant = "raster" if raster_ms >= max(stage[0], stage[1], stage[3]) else \
            max(("setup", stage[0]), ("bin_count", stage[1]), ("bin_dispatch", stage[3]), key=lambda t: t[1])[0]
        dom_ms = {"raster": raster_ms, "setup": stage[0], "bin_count": stage[1], "bin_dispatch": stage[3]}[dominant]
        achieved = ab[dominant] / (dom_ms * 1e-3) / 1e9
        traffic, issue_pct, traffic_src = None, None, None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath) and not split:
            with open(tpath) as f:
                tfile = json.load(f)
            captured = tfile.get(f"config{args.config}", {})
            # only a capture taken with the kernels as they are now describes this run
            if tfile.get("kernel_source_hash") == kernel_source_hash():
                traffic = captured.get(dominant)
                issue_pct = captured.get("issue_active_pct", {}).get(dominant)
                traffic_src = captured.get("source")
            else:
                traffic_src = "none: profiles/dram_traffic.json was captured with different kernel sources"
        counters = {k: stats[k] for k in ("visible_small", "visible_large", "bin_quads", "bin_tris", "low_bins",
                                          "high_bins", "promoted_bins", "fragments", "half_block_tris")}
        if depth is not None:
            counters["depth_complexity"] = depth  # fragments per covered pixel of this frame
        if split:
            counters["scope"] = "rank 0's bin range only (every rank counts its own bins)"
        line = {
            "metric": "exact_oit_frames_per_sec", "value": round(value, 3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "strong" if split else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "resolution": [width, height],
                       "input_triangles": tris_per_frame, "scale": args.scale,
                       "parallelism": ("bin-row split x%d (%s), P2P composite; rank 0 owns %s" %
                                       (world, split_kind + "; composite by " +
                                        ("a bin-row copy kernel" if args.composite == "copy" else "direct raster stores") +
                                        "; frame completion by " + ("P2P flags (no collective)" if use_flags else "NCCL all-reduce"),
                                        list(rows))) if split else "single GPU",
                       "l2": "256 MiB device memset between timed frames (untimed); the frame's own records "
                             "(~1 GB) exceed the 126 MB L2"},
            "mtris_per_sec": round(value * tris_per_frame / 1e6, 2),
            "ms_per_step_median": round(step_median, 4),
            "stage_ms": stage_ms, "rank_frame_ms": rank_frame_ms,
            "stage_ms_source": "a second pass over the same %d frames with a CUDA event after every stage; the timed "
                               "frames only carry the frame's first and last event (events between kernels disable "
                               "the launch overlap)" % kept,
            "counters": counters,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": round(achieved, 2), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes": int(ab[dominant]),
                         "kernel_ms": round(float(dom_ms), 4),
                         # from the committed ncu --set full capture of one frame of this workload (profiles/):
                         # DRAM bytes of the stage's kernels, and their SM issue-slot utilisation -- the raster
                         # kernels are bound by instruction issue, not by HBM
                         "traffic_source": traffic_src, "sm_issue_active_pct": issue_pct,
                         "per_stage_frac": {k: (round(v, 4) if v is not None else None) for k, v in fracs.items()}},
            "e2e": dict({"value": round(e2e_value, 3), "unit": "frames/s",
                         "h2d_bytes_per_step": int((len(inst) * 36 + 352) * (world if split else 1)),
                         "d2h_bytes_per_step": int(width * height * 4 + info.size * 4 * (world if split else 1))},
                        **e2e_extra),
            "sustained": sustained,
            # k_frame_begin, k_instance_select, k_quad_cull, k_tri_setup, k_bin_count, k_bin_scan, k_bin_dispatch,
            # k_raster_bins, k_block_sort, k_tie_runs, k_block_shade (+ k_info_out when LucidInfo is read back; the
            # split adds the frame gate / flag kernels of sync.cu)
            "gpu_launches": int(11 * args.steps),
            "clocks": clocks,
            "wall_s": round(wall, 3),
        }
        if inflight is not None:
            # value / ms_per_step are the pipelined run's; the classic loop (one handle, one frame after the other,
            # L2 flushed in between) stays beside it, and `sustained`, `stage_ms`, `rank_frame_ms` describe that loop
            line["frames_in_flight"] = inflight["frames_in_flight"]
            line["one_frame_at_a_time"] = {"value": round(1000.0 / serial_ms_per_step, 3), "unit": "frames/s",
                                           "ms_per_step": round(serial_ms_per_step, 4),
                                           "ms_per_step_median": round(step_median, 4), "sustained": sustained}
            line["sustained"] = inflight["sustained"]
            line["config"]["l2"] = ("no flush between the pipelined frames: a frame (a rank's share of it) reads several "
                                    "hundred MB of geometry and records, the L2 holds 126 MB; the one-frame-at-a-time "
                                    "loop flushes with a 256 MiB memset between timed frames (untimed)")
            line["config"]["parallelism"] += "; %d frames in flight (handles on separate streams)" % inflight["frames_in_flight"]
        if frames_in_flight_tried is not None:
            line["frames_in_flight"] = 1
            line["frames_in_flight_tried"] = frames_in_flight_tried
        if views is not None:
            line["views"] = views
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(scene, args)
        print(json.dumps(line), flush=True)
    if views is None:
        lane0.close()
        rig.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(scene, args, frames: int = 3):
    """The oracle (CPU restatement of the reference shaders) timed on the host cores: one warm-up frame, then
    `frames` frames of the same workload and camera (a bounded sample: the 4K scenes take seconds per frame)."""
    from lucid_b200 import api
    from oracle.binding import Oracle
    threads = os.cpu_count() or 1
    o = Oracle(scene["width"], scene["height"], 0, args.mvq or 4793490, threads=threads)
    o.set_scene(scene)
    cfg, inst, cols, rects = api.prepare_frame(scene)
    o.render(cfg, inst, cols, rects)
    times = []
    for _ in range(frames):
        t0 = time.perf_counter()
        o.render(cfg, inst, cols, rects)
        times.append(time.perf_counter() - t0)
    o.close()
    return {"value": round(1.0 / float(np.mean(times)), 4), "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{frames} full frames of the same workload and camera after one warm-up frame",
            "note": "CPU restatement of the reference shaders (OpenMP); Vulkan/lavapipe unavailable on this image"}


def run_reference(args):
    """The reference arm: the CPU restatement on all host threads, on rank 0 only.  Loads oracle/ and the
    host-only input preparation library -- never the kernels' library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lucid_b200 import api
    from oracle.binding import Oracle
    scene = make_scene(args.config, args.scale)
    threads = os.cpu_count() or 1
    o = Oracle(scene["width"], scene["height"], 0, args.mvq or 4793490, threads=threads)
    o.set_scene(scene)
    inst, cols, rects = api.build_instances(scene["draw_calls"], scene["materials"])

    def frame(step):
        cam = api.make_camera(view_camera(scene, step % 64), scene["width"], scene["height"])
        cfg = api.make_config(cam, len(inst), scene["background"])
        o.render(cfg, inst, cols, rects)

    # a frame of the 4K scenes takes the CPU seconds: the run is bounded to about two minutes of CPU work
    t0 = time.perf_counter()
    frame(0)
    first = time.perf_counter() - t0
    budget = 120.0
    warmup = max(0, min(args.warmup, int(0.2 * budget / max(first, 1e-3))) - 1)
    steps = max(1, min(args.steps, int(0.8 * budget / max(first, 1e-3))))
    for w in range(warmup):
        frame(1 + w)
    t0 = time.perf_counter()
    for k in range(steps):
        frame(1 + warmup + k)
    dt = time.perf_counter() - t0
    value = steps / dt
    stats = api.decode_stats(o.info, o.bin_count, o.width, o.height)
    sample = "every step is one full frame of the workload on all host threads"
    if steps != args.steps:
        sample += f"; bounded to {steps} timed steps after {warmup + 1} warm-up frames ({first:.1f} s per frame)"
    line = {
        "impl": "reference", "metric": "exact_oit_frames_per_sec", "value": round(value, 4), "unit": "frames/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warmup + 1,
        "ms_per_step": round(1000.0 * dt / steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "resolution": [scene["width"], scene["height"]],
                   "input_triangles": 2 * stats["input_quads"], "scale": args.scale},
        "cpu_baseline": {"value": round(value, 4), "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=DEFAULT_CONFIG)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--mode", default=None, choices=["views", "split"],
                    help="N>1: split (default) = bin ranges of one frame per rank; views = one orbit view per rank")
    ap.add_argument("--mvq", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-views", action="store_true", help="--mode split: skip the secondary views-sharded run")
    ap.add_argument("--sustained-seconds", type=float, default=1.0)
    ap.add_argument("--equal-rows", action="store_true", help="--mode split: equal row counts instead of cost-balanced")
    ap.add_argument("--split-rows", action="store_true", help="--mode split: whole bin rows, balanced on measured cost")
    ap.add_argument("--composite", default="stores", choices=["copy", "stores"],
                    help="--mode split: how the other ranks' strips reach rank 0's image")
    ap.add_argument("--trace-split", action="store_true", help="--mode split: per-rank step times (stderr)")
    ap.add_argument("--completion", default="flags", choices=["flags", "allreduce"],
                    help="--mode split: how rank 0 learns that every strip of a frame has landed")
    ap.add_argument("--balance-iters", type=int, default=7, help="--mode split: feedback steps of the range balancing")
    ap.add_argument("--balance-iters-in-flight", type=int, default=5,
                    help="--mode split: further feedback steps on the ranks' pipelined frame times")
    ap.add_argument("--frames-in-flight-single", type=int, default=2,
                    help="N = 1: renderer handles rendering alternate frames on their own streams (1 = off)")
    ap.add_argument("--frames-in-flight", type=int, default=3,
                    help="--mode split: renderer handles per rank rendering alternate frames on their own streams")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
