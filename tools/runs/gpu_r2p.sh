#!/bin/bash
tag=${1:-r2p}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "quadgen" > $out/${tag}_pytest_quadgen.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_quadgen.log
tail -3 $out/${tag}_pytest_quadgen.log
timeout 600 python tools/quadgen_bench.py --tris 10000000 > $out/${tag}_quadgen_grid.json 2> $out/${tag}_quadgen.err
timeout 600 python tools/quadgen_bench.py --tris 4000000 --mesh delaunay --cpu-tris 500000 > $out/${tag}_quadgen_delaunay.json 2>> $out/${tag}_quadgen.err
cat $out/${tag}_quadgen_grid.json $out/${tag}_quadgen_delaunay.json; tail -3 $out/${tag}_quadgen.err
timeout 900 python tools/inflight_probe.py 3 60 > $out/${tag}_inflight_probe.txt 2>&1
cat $out/${tag}_inflight_probe.txt | tail -8
