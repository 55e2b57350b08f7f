#!/bin/bash
# opaque pre-pass on a B200: its GPU tests, then stage times of configs 1 and 3 without and with the option
tag=${1:-pp}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "prepass" > $out/${tag}_pytest_prepass.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_prepass.log
tail -5 $out/${tag}_pytest_prepass.log
timeout 600 python tools/gpu_probe.py 1 3 > $out/${tag}_probe_plain.txt 2>&1
timeout 600 python tools/gpu_probe.py 1 3 --opts=0x100 > $out/${tag}_probe_prepass.txt 2>&1
grep -E "^==|stage_ms" $out/${tag}_probe_plain.txt $out/${tag}_probe_prepass.txt
