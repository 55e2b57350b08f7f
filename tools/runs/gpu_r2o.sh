#!/bin/bash
# quad pairing tests + throughput after the chunk-skipping change; per-kernel launch list of a split played on one GPU
tag=${1:-r2o}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "quadgen" > $out/${tag}_pytest_quadgen.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_quadgen.log
tail -5 $out/${tag}_pytest_quadgen.log
timeout 600 python tools/quadgen_bench.py --tris 10000000 > $out/${tag}_quadgen_grid.json 2> $out/${tag}_quadgen.err
timeout 600 python tools/quadgen_bench.py --tris 4000000 --mesh delaunay --cpu-tris 500000 > $out/${tag}_quadgen_delaunay.json 2>> $out/${tag}_quadgen.err
cat $out/${tag}_quadgen_grid.json $out/${tag}_quadgen_delaunay.json; tail -3 $out/${tag}_quadgen.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_split_launches.csv python tools/split_probe.py 3 8 --cull > $out/${tag}_split_probe_under_ncu.txt 2>&1
tail -3 $out/${tag}_split_probe_under_ncu.txt
wc -l $out/${tag}_split_launches.csv
