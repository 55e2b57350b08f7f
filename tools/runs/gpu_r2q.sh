#!/bin/bash
# instance selection (k_instance_select): whole GPU suite, split played on one GPU with --cull, N=1 bench line
tag=${1:-r2q}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
timeout 600 python tools/split_probe.py 3 8 --cull > $out/${tag}_split_probe.txt 2>&1
cat $out/${tag}_split_probe.txt | cut -c1-420
timeout 600 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_config3.json 2> $out/${tag}_bench.err
cat $out/${tag}_bench_config3.json | cut -c1-1500; tail -3 $out/${tag}_bench.err
timeout 600 python tools/quadgen_bench.py --tris 10000000 > $out/${tag}_quadgen_grid.json 2> $out/${tag}_quadgen.err
cat $out/${tag}_quadgen_grid.json
