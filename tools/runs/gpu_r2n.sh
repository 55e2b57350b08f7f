#!/bin/bash
# opaque pre-pass + quad pairing on a B200: their GPU tests, stage times without / with the pre-pass, pairing throughput
tag=${1:-r2n}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "prepass or quadgen" > $out/${tag}_pytest_prepass_quadgen.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_prepass_quadgen.log
tail -15 $out/${tag}_pytest_prepass_quadgen.log
timeout 600 python tools/quadgen_bench.py --tris 10000000 > $out/${tag}_quadgen_grid.json 2> $out/${tag}_quadgen.err
timeout 600 python tools/quadgen_bench.py --tris 4000000 --mesh delaunay --cpu-tris 500000 > $out/${tag}_quadgen_delaunay.json 2>> $out/${tag}_quadgen.err
cat $out/${tag}_quadgen_grid.json $out/${tag}_quadgen_delaunay.json; tail -3 $out/${tag}_quadgen.err
timeout 600 python tools/gpu_probe.py 1 3 > $out/${tag}_probe_plain.txt 2>&1
timeout 600 python tools/gpu_probe.py 1 3 --opts=0x100 > $out/${tag}_probe_prepass.txt 2>&1
grep -E "^==|stage_ms" $out/${tag}_probe_plain.txt $out/${tag}_probe_prepass.txt
