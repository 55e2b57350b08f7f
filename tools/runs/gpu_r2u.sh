#!/bin/bash
# debug-raster variant tests; frames in flight on one GPU playing ranks 0 / 3 / 7 of 8: handle counts, shade grid sizes
tag=${1:-r2u}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "debug_raster or owned_bins" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log; tail -3 $out/${tag}_pytest.log
PROBE_HANDLES=5 timeout 900 python tools/inflight_probe.py 3 40 > $out/${tag}_inflight_5.txt 2>&1; tail -4 $out/${tag}_inflight_5.txt
for g in 4 5; do
LUCID_SHADE_GRID_CTAS=$g PROBE_HANDLES=3 timeout 900 python tools/inflight_probe.py 3 40 > $out/${tag}_inflight_grid$g.txt 2>&1; echo "shade grid $g"; tail -4 $out/${tag}_inflight_grid$g.txt
done
