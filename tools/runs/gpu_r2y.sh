#!/bin/bash
# write-pass unroll of k_block_sort: parity of the raster tests, stage times, split played on one GPU; ncu capture of config 3
tag=${1:-r2y}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log; tail -3 $out/${tag}_pytest_gpu.log
timeout 600 python tools/gpu_probe.py 1 2 3 > $out/${tag}_probe_stage_times.txt 2>&1
grep -E "^==|stage_ms" $out/${tag}_probe_stage_times.txt | cut -c1-330
timeout 600 python tools/split_probe.py 3 8 --cull > $out/${tag}_split_probe.txt 2>&1
cut -c1-330 $out/${tag}_split_probe.txt
bash tools/gpu_round2.sh $tag "ncu" "3" > $out/${tag}_round2.log 2>&1
head -60 $out/${tag}_ncu_full_config3.txt
