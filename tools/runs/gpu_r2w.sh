#!/bin/bash
# compact lists on configs 2, 3 (default pool); CTA size of k_raster_bins on the ranks of an 8-way split played on one GPU
tag=${1:-r2w}
out=gpurun_out
mkdir -p $out
timeout 600 python tools/gpu_probe.py 2 3 --compact > $out/${tag}_probe_compact.txt 2>&1
grep -E "^==|stage_ms|rror" $out/${tag}_probe_compact.txt | cut -c1-400
for t in 256 512 1024; do
  LUCID_RASTER_BINS_THREADS=$t timeout 600 python tools/split_probe.py 3 8 --cull > $out/${tag}_split_probe_bins$t.txt 2>&1
  echo "k_raster_bins threads $t"; grep -E "^rank|^sum" $out/${tag}_split_probe_bins$t.txt | sed -E "s/.*'lists': ([0-9.]+).*'frame': ([0-9.]+).*/lists \1 frame \2/"
done
