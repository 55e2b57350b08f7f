#!/bin/bash
# cooperative ranking of long tie runs: stress test, whole GPU suite, stage times
tag=${1:-r3j}
out=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "depth_ties" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log; tail -3 $out/${tag}_pytest_gpu.log
timeout 600 python tools/gpu_probe.py 1 2 3 > $out/${tag}_probe_stage_times.txt 2>&1
grep -E "^==|stage_ms" $out/${tag}_probe_stage_times.txt | cut -c1-330
