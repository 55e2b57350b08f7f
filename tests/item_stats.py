"""How full would the chunked shading loop of k_raster_blocks be?  Counted on the CPU checker's block
lists (no GPU needed; test infrastructure, like everything that loads oracle/):
   python tests/item_stats.py [config ...] > profiles/<tag>_item_statistics.txt"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

for ci in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    sc = scenes.get_config(ci)
    o = Oracle(sc["width"], sc["height"], 0, 4793490, threads=os.cpu_count() or 1)
    o.lib.oracle_set_item_stats.argtypes = [C.c_void_p, C.c_int]
    o.lib.oracle_read_item_stats.argtypes = [C.c_void_p, C.c_void_p]
    o.lib.oracle_set_item_stats(o.h, 1)
    o.set_scene(sc)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    o.render(cfg, inst, cols, rects)
    st = (C.c_ulonglong * 24)()
    o.lib.oracle_read_item_stats(o.h, st)
    st = list(st)
    lists, entries, samples = st[0], st[1], st[2]
    print(f"== config {ci} ({sc['name']}): {lists} half-block lists, {entries} entries ({entries / max(lists, 1):.1f} per list), "
          f"{samples} samples ({samples / max(entries, 1):.2f} per entry)")
    for w, base in ((32, 3), (64, 6)):
        chunks, rounds, iters = st[base], st[base + 1], st[base + 2]
        print(f"   chunks of {w} entries: {chunks} chunks, {samples / max(chunks, 1):.1f} samples per chunk; sample-parallel "
              f"shading {rounds} rounds -> {100 * samples / max(32 * rounds, 1):.0f}% of lanes busy; pixel-parallel reduce "
              f"{iters} iterations -> {100 * samples / max(32 * iters, 1):.0f}% of lanes busy")
    names = ["> 384", "161-384", "65-160", "25-64", "<= 24"]
    for k, nm in enumerate(names):
        print(f"   size class {nm:8s}: {st[9 + k]:8d} lists ({100 * st[9 + k] / max(lists, 1):5.1f}%), "
              f"{st[14 + k]:9d} entries ({100 * st[14 + k] / max(entries, 1):5.1f}%)")
    o.close()
