#!/usr/bin/env python
"""bench.py -- exact-OIT frames/s of the B200 rasteriser on BASELINE.json's synthetic configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config I] [--impl ours|reference]

A "step" is one frame: quad setup -> binning -> raster low/high -> image, for one view of the
workload scene.  At N=1 the workload is BASELINE.json configs[1] (1M-triangle meshlet scene,
1920x1080).  With N>1 every rank renders its own view of an orbit around the same scene per step
(views sharded per GPU, no data-path collective: "scaling": "weak"); `--mode split` instead splits
the bin rows of ONE frame over the ranks and stores every strip straight into rank 0's image over
NVLink (strong scaling, reported for config 3 in DESIGN.md).

`value` is frames/s with geometry resident in HBM (per-frame config + instance upload included);
`e2e` is the same frame through lucid_render() with host buffers in and the RGBA8 image plus
LucidInfo copied back to the host inside the timed region.  `--impl reference` times the CPU
restatement of the reference shaders (oracle/) with all host threads: the reference's own Vulkan
path cannot run on this image (no ICD, no shaderc).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    0: "config0: 100k-triangle quad soup, 50% alpha=0.5, 1280x720",
    1: "config1: 1M-triangle meshlet scene (489 patches of 32x32 quads), mixed opaque/transparent, 1920x1080",
    2: "config2: hairball-like dense overlap, 5M triangles, 3840x2160",
    3: "config3: 10M-triangle architecture scene, textured atlas shading, 3840x2160",
    4: "config4: 64-view orbit of the 1M-triangle scene, 1920x1080",
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~2 ms while the timed region runs
    (the region lasts tens of milliseconds, too short for `nvidia-smi -lms`)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reason_bits = [], 0
        self.stop_flag = False
        self.thread = None
        self.handle = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() \
                else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.handle = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        reasons = sorted(n for n, bit in self.REASONS.items() if self.reason_bits & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm)}


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

    scene = make_scene(args.config, args.scale)
    width, height = scene["width"], scene["height"]
    # a stream of our own, made torch's current stream: the library launches every kernel on the stream it
    # is handed (handle 0 would make it create a private one, invisible to torch.cuda.Event), so the
    # flush, the timing events and the kernels are all ordered on this one stream
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    split = args.mode == "split" and world > 1
    nby = (height + 31) // 32
    r = api.LucidRenderer(width, height, 0, args.mvq, device=local_rank, stream=stream.cuda_stream)
    r.set_scene(scene)
    inst, cols, rects = api.build_instances(scene["draw_calls"], scene["materials"])
    rows = None
    if split:
        # ownership from the measured raster cost of a full calibration frame (what an application takes
        # from the previous frame); rank 0's measurement is used by every rank.  Default: contiguous
        # row-major bin ranges of equal cost (a heavy bin row may be shared by two ranks); --split-rows
        # keeps whole bin rows, --equal-rows equal row counts.
        from lucid_b200 import multigpu
        cam0 = api.make_camera(view_camera(scene, 0), width, height)
        for _ in range(2):
            r.render(api.make_config(cam0, len(inst), scene["background"]), inst, cols, rects)
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
            # rescaled by the time that rank actually needed for them (everything but the replicated
            # setup), and the ranges are cut again
            for _ in range(args.balance_iters):
                for k in range(3):
                    r.render(api.make_config(cam0, len(inst), scene["background"]), inst, cols, rects,
                             flags=api.RENDER_ASYNC | api.RENDER_SKIP_INFO)
                mine = float(np.median([r.stage_times(i)[1:7].sum() for i in range(2)]))
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

    def config_for(step):
        view = 0 if split else (step * world + rank) % 64
        cam = api.make_camera(view_camera(scene, view), width, height)
        return api.make_config(cam, len(inst), scene["background"])

    # composite target for the bin-row split: every rank stores into rank 0's image over NVLink
    peer_ptr = None
    if split:
        handle = [r.ipc_export_image() if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            peer_ptr = r.ipc_open_image(handle[0])

    def render(step, **kw):
        if peer_ptr is not None and args.composite == "stores":
            # the raster kernels store their pixels straight into rank 0's image
            r.render(config_for(step), inst, cols, rects, out_device_ptr=peer_ptr, out_pitch=width * 4, **kw)
        else:
            r.render(config_for(step), inst, cols, rects, **kw)
            if peer_ptr is not None:  # own image first, then the owned bins as whole 128-byte rows
                r.composite_to(peer_ptr, width * 4)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    frame_token = torch.zeros(1, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # timed frames carry only the frame's first and last event: per-stage events between the kernels would
    # keep each launch from overlapping its predecessor's tail (programmatic dependent launch)
    plain = api.RENDER_ASYNC | api.RENDER_SKIP_INFO | api.RENDER_NO_STAGE_TIMES
    for w in range(args.warmup):  # the same sequence as a timed step, completion all-reduce included
        render(w, flags=plain)
        if split:
            dist.all_reduce(frame_token)
    barrier()

    # NVML polling takes driver locks: only the rank that prints the line samples its GPU
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # --trace-split: an extra event between the frame and the completion all-reduce of every step
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)] if (split and args.trace_split) else None
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_wall = time.time()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # L2 flush between timed iterations (untimed)
        starts[k].record(stream)
        render(args.warmup + k, flags=plain)
        if split:  # the frame is complete when every rank's strip has landed in rank 0's image
            if mids is not None:
                mids[k].record(stream)
            dist.all_reduce(frame_token)
        stops[k].record(stream)
    barrier()
    wall = time.time() - t_wall
    kept = min(args.steps, 64)
    frame_ms = float(np.mean([r.stage_times(i)[7] for i in range(kept)]))  # library's own first->last event
    rank_frame_ms = None
    if dist is not None:  # every rank's own frame time (the split's balance; a step ends with the slowest)
        t = torch.zeros(world, device="cuda", dtype=torch.float64)
        t[rank] = frame_ms
        dist.all_reduce(t)
        rank_frame_ms = [round(float(x), 4) for x in t.cpu().numpy()]
    step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, stops)], np.float64)
    if mids is not None:  # where a step's time goes on this rank: submission + frame, then waiting for the others
        a = np.median([s.elapsed_time(m) for s, m in zip(starts, mids)])
        b = np.median([m.elapsed_time(e) for m, e in zip(mids, stops)])
        lib = np.median([r.stage_times(i)[7] for i in range(min(args.steps, 64))])
        sys.stderr.write(f"[trace-split] rank {rank}: start->frame end {a:.4f} ms (library frame {lib:.4f} ms), "
                         f"frame end->all-reduce done {b:.4f} ms, step median {np.median(step_ms):.4f} ms; steps "
                         f"{[round(float(x), 3) for x in step_ms]}\n")
    total_ms = torch.tensor([float(step_ms.sum())], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    # per-stage CUDA-event times: the same frames again, same L2 flush, this time with an event after every
    # stage (recorded on the stream the kernels are launched on; the library keeps its last 64 frames)
    for k in range(kept):
        flush.fill_(k & 0xFF)
        render(args.warmup + k, flags=api.RENDER_ASYNC | api.RENDER_SKIP_INFO)
    barrier()
    stage = np.mean([r.stage_times(i).astype(np.float64) for i in range(kept)], axis=0)

    # end to end through the C ABI: host instance arrays in (the library stages them through pinned
    # memory), RGBA8 image + LucidInfo read back into pinned host memory every frame.  Frames are
    # submitted asynchronously: the copy-out of frame n overlaps the rendering of frame n+1.
    host_imgs = [torch.empty((height, width), dtype=torch.int32).pin_memory() for _ in range(2)]
    e2e_steps = max(args.steps, 3)

    def e2e_frame(k):
        if split:
            # every rank rasterises its bin rows into rank 0's image, then rank 0 reads it back
            render(k, flags=api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES)
            dist.all_reduce(frame_token)
            if rank == 0:
                torch.cuda.current_stream().synchronize()
                r.read_image_into(host_imgs[k & 1].data_ptr())
        else:
            r.render(config_for(k), inst, cols, rects, out=host_imgs[k & 1].data_ptr(),
                     flags=api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES)

    for w in range(3):
        e2e_frame(w)
    r.wait()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_frame(k)
    r.wait()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    clocks = sampler.stop()

    # counters of one frame for the roofline arithmetic
    r.render(config_for(0), inst, cols, rects)
    info = r.read_info()
    stats = api.decode_stats(info, r.bin_count, width, height)

    frames_per_step = 1 if split else world
    tris_per_frame = 2 * stats["input_quads"]
    ms_per_step = total_ms / args.steps
    value = frames_per_step * 1000.0 / ms_per_step
    e2e_value = frames_per_step * e2e_steps / e2e_s

    if rank == 0:
        peak, peak_src = load_peaks()
        ab = algorithmic_bytes(stats, scene, width, height)
        # LOW and HIGH bins share the two raster kernels (block lists, then block sort + shading)
        stage_names = ["setup", "bin_count", "bin_scan", "bin_dispatch", "raster_lists", "raster_shade", "finish"]
        stage_ms = {n: round(float(stage[i]), 4) for i, n in enumerate(stage_names)}
        stage_ms["frame_with_stage_events"] = round(float(stage[7]), 4)
        stage_ms["frame"] = round(frame_ms, 4)  # timed frames: first launch -> last kernel done, the library's own events
        # the library's events and ours are on one stream: our bracket can only be the wider one
        assert ms_per_step >= 0.98 * frame_ms, (ms_per_step, frame_ms)
        raster_ms = float(stage[4] + stage[5])
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
        if os.path.exists(tpath):
            with open(tpath) as f:
                captured = json.load(f).get(f"config{args.config}", {})
            traffic = captured.get(dominant)
            issue_pct = captured.get("issue_active_pct", {}).get(dominant)
            traffic_src = captured.get("source")
        line = {
            "metric": "exact_oit_frames_per_sec", "value": round(value, 3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "strong" if split else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "resolution": [width, height],
                       "input_triangles": tris_per_frame, "scale": args.scale,
                       "parallelism": ("bin-row split x%d (%s), P2P composite; rank 0 owns %s" %
                                       (world, split_kind + "; composite by " +
                                        ("a bin-row copy kernel" if args.composite == "copy" else "direct raster stores"),
                                        list(rows))) if split else
                       ("views sharded x%d" % world if world > 1 else "single GPU"),
                       "l2": "256 MiB device memset between timed frames (untimed)"},
            "mtris_per_sec": round(value * tris_per_frame / 1e6, 2),
            "stage_ms": stage_ms, "rank_frame_ms": rank_frame_ms,
            "stage_ms_source": "a second pass over the same %d frames with a CUDA event after every stage; the timed "
                               "frames only carry the frame's first and last event (events between kernels disable "
                               "the launch overlap)" % kept,
            "counters": {k: stats[k] for k in ("visible_small", "visible_large", "bin_quads", "bin_tris", "low_bins",
                                               "high_bins", "promoted_bins", "fragments", "half_block_tris")},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": round(achieved, 2), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes": int(ab[dominant]),
                         # from the committed ncu --set full capture of one frame of this workload (profiles/):
                         # DRAM bytes of the stage's kernels, and their SM issue-slot utilisation -- the raster
                         # kernels are bound by instruction issue, not by HBM
                         "traffic_source": traffic_src, "sm_issue_active_pct": issue_pct,
                         "per_stage_frac": {k: (round(v, 4) if v is not None else None) for k, v in fracs.items()}},
            "e2e": {"value": round(e2e_value, 3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(len(inst) * 36 + 352),
                    "d2h_bytes_per_step": int(width * height * 4 + info.size * 4)},
            # k_frame_begin, k_quad_cull, k_tri_setup, k_bin_count, k_bin_scan, k_bin_dispatch, k_raster_bins,
            # k_raster_blocks (+ k_info_out when LucidInfo is read back)
            "gpu_launches": int(8 * args.steps),
            "clocks": clocks,
            "wall_s": round(wall, 3),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(scene, args)
        print(json.dumps(line), flush=True)
    if peer_ptr is not None:
        r.ipc_close_image(peer_ptr)
    r.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(scene, args, frames: int = 1):
    """The oracle (CPU restatement of the reference shaders) timed on the host cores."""
    from lucid_b200 import api
    from oracle.binding import Oracle
    threads = os.cpu_count() or 1
    o = Oracle(scene["width"], scene["height"], 0, args.mvq or 4793490, threads=threads)
    o.set_scene(scene)
    cfg, inst, cols, rects = api.prepare_frame(scene)
    times = []
    for _ in range(frames):
        t0 = time.perf_counter()
        o.render(cfg, inst, cols, rects)
        times.append(time.perf_counter() - t0)
    o.close()
    return {"value": round(1.0 / float(np.mean(times)), 4), "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{frames} full frame(s) of the same workload and camera",
            "note": "CPU restatement of the reference shaders (OpenMP); Vulkan/lavapipe unavailable on this image"}


def run_reference(args):
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

    for w in range(args.warmup):
        frame(w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        frame(args.warmup + k)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    stats = api.decode_stats(o.info, o.bin_count, o.width, o.height)
    line = {
        "impl": "reference", "metric": "exact_oit_frames_per_sec", "value": round(value, 4), "unit": "frames/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1000.0 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "resolution": [scene["width"], scene["height"]],
                   "input_triangles": 2 * stats["input_quads"], "scale": args.scale},
        "cpu_baseline": {"value": round(value, 4), "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "every step is one full frame of the workload (one orbit view per step)"},
        "e2e": {"value": round(value, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--mode", default="views", choices=["views", "split"])
    ap.add_argument("--mvq", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--equal-rows", action="store_true", help="--mode split: equal row counts instead of cost-balanced")
    ap.add_argument("--split-rows", action="store_true", help="--mode split: whole bin rows, balanced on measured cost")
    ap.add_argument("--composite", default="stores", choices=["copy", "stores"],
                    help="--mode split: how the other ranks' strips reach rank 0's image")
    ap.add_argument("--trace-split", action="store_true",
                    help="--mode split: per-rank split of a step into frame and completion wait (stderr)")
    ap.add_argument("--balance-iters", type=int, default=3, help="--mode split: feedback steps of the range balancing")
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
