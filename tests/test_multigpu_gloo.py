"""N>1 host logic on CPU: two gloo ranks split the bin rows of one frame, each produces its strip
(with the oracle standing in for the GPU), and the strips are composited on rank 0."""
import os
import socket

import numpy as np
import pytest

from lucid_b200 import multigpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    from lucid_b200 import api
    from tests import parity_util as pu

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = pu.small_scenes()["soup_close"]
        h, w = sc["height"], sc["width"]
        nby = (h + 31) // 32
        all_rows = multigpu.split_bin_rows(nby, world)
        rows = all_rows[rank]
        part = pu.run_oracle(sc, threads=2, bin_rows=rows)
        y0, y1 = multigpu.strip_pixel_rows(rows, h)
        strip = torch.from_numpy(part.read_image()[y0:y1].view(np.int32).copy())
        full = multigpu.composite_gather(dist, strip, rows, all_rows, h, w, dst=0)
        summed = multigpu.reduce_info(dist, part.info, part.bin_count)
        views = multigpu.views_for_rank(7, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, views)
        if rank == 0:
            ref = pu.run_oracle(sc, threads=2)
            ok_img = np.array_equal(full.numpy().view(np.uint32), ref.read_image())
            # statistics, rejections, per-bin counts, offsets and level counts of the composed frame equal the
            # single-device frame's; the LOW / HIGH lists as sets
            hs, cs = api.split_info(summed, ref.bin_count)
            hr, cr = api.split_info(ref.info, ref.bin_count)
            ok_frag = (np.array_equal(hs[60:63], hr[60:63]) and hs[0] == hr[0] and np.array_equal(hs[32:36], hr[32:36])
                       and np.array_equal(cs[:6], cr[:6]) and np.array_equal(hs[5:10], hr[5:10])
                       and set(cs[7][:hs[7]]) == set(cr[7][:hr[7]]) and set(cs[9][:hs[9]]) == set(cr[9][:hr[9]]))
            ok_views = sorted(v for g in gathered for v in g) == list(range(7))
            with open(out_path, "w") as f:
                f.write(f"{int(ok_img)} {int(ok_frag)} {int(ok_views)}")
    finally:
        dist.destroy_process_group()


def test_two_rank_bin_row_split_composite(tmp_path):
    import torch.multiprocessing as mp

    out = tmp_path / "result.txt"
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "1 1 1"


def test_split_bin_rows_properties():
    for nby in (1, 7, 23, 34, 68):
        for world in (1, 2, 3, 4, 8):
            parts = multigpu.split_bin_rows(nby, world)
            assert parts[0][0] == 0 and parts[-1][1] == nby
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            if nby >= world:
                assert all(e > b for b, e in parts)
    # weighted split puts the heavy rows in separate ranks
    w = np.zeros(68)
    w[30:34] = 100.0
    parts = multigpu.split_bin_rows(68, 4, w)
    assert len({next(i for i, (b, e) in enumerate(parts) if b <= r < e) for r in range(30, 34)}) >= 3
    with pytest.raises(ValueError):
        multigpu.split_bin_rows(10, 2, np.ones(3))


def test_split_bins_properties():
    """Row-major bin ranges of equal cost: contiguous, complete, non-empty, balanced up to one bin."""
    rng = np.random.default_rng(3)
    for bins, world in ((2040, 2), (8160, 8), (920, 3), (7, 8)):
        w = rng.random(bins) ** 4 * 1000.0
        w[rng.random(bins) < 0.4] = 0.0  # empty bins cost nothing
        parts = multigpu.split_bins(bins, world, w)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == bins
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        if bins >= world:
            assert all(hi > lo for lo, hi in parts)
            sums = np.array([w[lo:hi].sum() for lo, hi in parts])
            assert sums.max() <= w.sum() / world + w.max() + 1e-6
    with pytest.raises(ValueError):
        multigpu.split_bins(10, 2, np.ones(9))
