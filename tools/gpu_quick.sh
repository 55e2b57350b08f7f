#!/bin/bash
# Short iteration pass on a B200 box: parity suite, stage times of the four single-GPU configs and
# (optionally) a full capture of k_raster_blocks.   bash tools/gpu_quick.sh <tag> [capture configs]
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 900 python tools/gpu_probe.py 0 1 2 3 > $out/${tag}_probe_stage_times.txt 2>&1
grep -E "^==|stage_ms" $out/${tag}_probe_stage_times.txt
for c in $2; do
  timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:k_raster_blocks --launch-skip 1 -c 1 \
    -o $out/${tag}_blocks_config$c -f python tools/profile_frame.py $c 2 > $out/${tag}_blocks_config$c.log 2>&1
  python tools/ncu_summary.py $out/${tag}_blocks_config$c.ncu-rep > $out/${tag}_ncu_blocks_config$c.txt 2>&1
done
