"""Host-side logic of the multi-GPU modes: bin-row partitioning of one frame with
replicated geometry, strip composite, and per-view sharding of multi-view batches.

One process per GPU (torch.distributed).  The fast composite needs no collective at all: rank 0
exports its image with CUDA IPC and every other rank's raster kernels store their strip straight
into it over NVLink (LucidRenderer.render(out_device_ptr=...)).  composite_gather() is the portable
path (NCCL on GPUs, gloo in the CPU tests): strips are disjoint, so it is a plain gather.
"""
from __future__ import annotations

import numpy as np

BIN_SIZE = 32


def split_bin_rows(bin_count_y: int, world_size: int, weights=None) -> list[tuple[int, int]]:
    """Contiguous bin-row ranges [begin, end) per rank.  With weights (e.g. last frame's fragments
    per bin row) the boundaries equalise the summed weight; every rank gets at least one row when
    bin_count_y >= world_size."""
    if world_size <= 0:
        raise ValueError("world_size")
    if weights is None:
        bounds = [round(i * bin_count_y / world_size) for i in range(world_size + 1)]
    else:
        w = np.asarray(weights, np.float64)
        if w.shape != (bin_count_y,):
            raise ValueError("weights must have one entry per bin row")
        c = np.concatenate([[0.0], np.cumsum(w + 1e-9)])
        targets = c[-1] * np.arange(1, world_size) / world_size
        inner = np.searchsorted(c, targets, side="left")
        bounds = [0] + [int(b) for b in inner] + [bin_count_y]
    # monotone, non-empty where possible
    for i in range(1, world_size):
        lo = bounds[i - 1] + (1 if bin_count_y >= world_size else 0)
        hi = bin_count_y - ((world_size - i) if bin_count_y >= world_size else 0)
        bounds[i] = min(max(bounds[i], lo), hi)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def split_bins(bin_count: int, world_size: int, weights) -> list[tuple[int, int]]:
    """Contiguous row-major bin ranges [begin, end) per rank with equal summed weight (per-bin raster
    cost of the previous frame).  Finer than split_bin_rows: one heavy bin row can be shared."""
    w = np.asarray(weights, np.float64)
    if w.shape != (bin_count,):
        raise ValueError("weights must have one entry per bin")
    c = np.concatenate([[0.0], np.cumsum(w + 1e-9)])
    targets = c[-1] * np.arange(1, world_size) / world_size
    bounds = [0] + [int(b) for b in np.searchsorted(c, targets, side="left")] + [bin_count]
    for i in range(1, world_size):
        lo = bounds[i - 1] + (1 if bin_count >= world_size else 0)
        hi = bin_count - ((world_size - i) if bin_count >= world_size else 0)
        bounds[i] = min(max(bounds[i], lo), hi)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def strip_pixel_rows(rows: tuple[int, int], height: int) -> tuple[int, int]:
    return min(rows[0] * BIN_SIZE, height), min(rows[1] * BIN_SIZE, height)


def views_for_rank(num_views: int, rank: int, world_size: int) -> list[int]:
    """Multi-view batches shard per view (north star): view v goes to rank v % world_size."""
    return list(range(rank, num_views, world_size))


def composite_gather(dist, strip, rows, all_rows, height, width, dst: int = 0):
    """Gathers disjoint image strips on rank dst.  strip: torch int32 tensor [strip_h, width] holding
    this rank's pixel rows; returns the full [height, width] tensor on dst (None elsewhere)."""
    import torch

    world = dist.get_world_size()
    rank = dist.get_rank()
    max_h = max(strip_pixel_rows(r, height)[1] - strip_pixel_rows(r, height)[0] for r in all_rows)
    padded = torch.zeros((max_h, width), dtype=strip.dtype, device=strip.device)
    padded[: strip.shape[0]] = strip
    if rank == dst:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, parts, dst=dst)
        full = torch.empty((height, width), dtype=strip.dtype, device=strip.device)
        for r, part in zip(all_rows, parts):
            y0, y1 = strip_pixel_rows(r, height)
            full[y0:y1] = part[: y1 - y0]
        return full
    dist.gather(padded, None, dst=dst)
    return None


def reduce_info(dist, info, bin_count: int):
    """LucidInfo + per-bin arrays of the whole frame from the ranks of a bin split (every rank passes its own
    lucid_read_info words).  Only what is additive is summed -- the statistics and the per-bin counts, which are
    disjoint because a rank only counts the bins it owns; the rest is rebuilt or taken as is:
      * num_input_quads and the rejection counters are identical on every rank (setup is replicated): maximum;
      * num_visible_quads / num_counted_quads are per rank (a quad is visible on every rank whose rows it
        touches) and are returned as the per-rank maximum, NOT a frame total;
      * bin offsets are the exclusive prefix sums of the summed counts (what bin_categorizer computes,
        bin_categorizer.glsl:22-89), the *_OFFSETS_TEMP arrays offsets + counts (their value after dispatch);
      * LOW / HIGH bin lists are the sorted unions of the ranks' lists, EMPTY the remainder.
    Returns the words as a uint32 array of the same length."""
    import torch

    words = np.asarray(info).astype(np.int64)
    head_n = 1152
    assert words.size >= head_n + 10 * bin_count
    head, counts = words[:head_n].copy(), words[head_n:head_n + 10 * bin_count].reshape(10, bin_count).copy()

    def allred(a, op):
        t = torch.from_numpy(np.array(a, dtype=np.int64, copy=True))  # a copy: all_reduce works in place
        dist.all_reduce(t, op=op)
        return t.numpy()

    SUM, MAX = dist.ReduceOp.SUM, dist.ReduceOp.MAX
    out_head = allred(head, MAX)  # identical or per-rank words: maximum
    out_head[60:64] = allred(head[60:64], SUM)  # stats: fragments, half-block triangles, invalid pixels
    out_head[1088:1090] = allred(head[1088:1090], MAX)  # temp[0] dropped quads (replicated), temp[1] list overflow
    quad_counts, tri_counts = allred(counts[0], SUM), allred(counts[3], SUM)
    # level membership: a rank lists only bins it owns; one-hot per bin, summed
    n_low, n_high = int(head[7]), int(head[9])
    member = np.zeros((2, bin_count), np.int64)
    member[0, counts[7][:n_low]] = 1
    member[1, counts[9][:n_high]] = 1  # includes bins promoted from LOW (they stay in the LOW list too)
    member = allred(member, SUM)
    low, high = np.nonzero(member[0])[0], np.nonzero(member[1])[0]
    out = np.zeros((10, bin_count), np.int64)
    out[0], out[3] = quad_counts, tri_counts
    out[1] = np.concatenate([[0], np.cumsum(quad_counts)[:-1]])
    out[4] = np.concatenate([[0], np.cumsum(tri_counts)[:-1]])
    out[2], out[5] = out[1] + out[0], out[4] + out[3]
    out[7][:low.size], out[9][:high.size] = low, high
    promoted = int(np.count_nonzero(member[0] * member[1]))
    out_head[5:10] = [bin_count - low.size - (high.size - promoted), 0, low.size, 0, high.size]
    return np.concatenate([out_head, out.reshape(-1)]).astype(np.uint32)


class SharedHostImages:
    """Host images shared by the processes of a split (POSIX shared memory under /dev/shm), page-locked in every
    process so that each device copies its own strip of a frame straight into them over its own PCIe link
    (LucidRenderer.render(out=..., flags=RENDER_OWNED_BINS_ONLY)): the frame is gathered in host memory, no device
    gathers it first.  `count` images of height x width RGBA8; rank 0 creates (and unlinks on close), the others attach."""

    def __init__(self, name: str, width: int, height: int, count: int, create: bool):
        import mmap
        import os
        self.path = "/dev/shm/" + name
        self.width, self.height, self.count = width, height, count
        self.nbytes = width * height * 4 * count
        self._owner = create
        fd = os.open(self.path, os.O_RDWR | (os.O_CREAT if create else 0), 0o600)
        try:
            if create:
                os.ftruncate(fd, self.nbytes)
            self._mm = mmap.mmap(fd, self.nbytes)
        finally:
            os.close(fd)
        self.array = np.frombuffer(self._mm, np.uint32).reshape(count, height, width)
        self._registered = False

    def pin(self):
        """cudaHostRegister in this process (needs torch with CUDA); copies into unpinned memory still work, slowly."""
        import torch
        rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, self.nbytes, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed: {rc}")
        self._registered = True

    def pointer(self, index: int) -> int:
        return self.array.ctypes.data + index * self.width * self.height * 4

    def close(self):
        import os
        if self._registered:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            self._registered = False
        self.array = None
        try:
            self._mm.close()
        except BufferError:
            pass
        if self._owner and os.path.exists(self.path):
            os.unlink(self.path)
