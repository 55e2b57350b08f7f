#!/bin/bash
# Bin-row split pass (run under gpurun --gpus N): balanced vs equal rows on configs 3 and 2, views-sharded
# bench, and the composite == single-GPU check.   bash tools/gpu_split.sh <tag> <N>
tag=${1:-r1}
n=${2:-2}
out=gpurun_out
mkdir -p $out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
for c in 3 2; do
  run --config $c --mode split --steps 10 > $out/${tag}_bench_split_config${c}_n$n.json 2> $out/${tag}_split_config${c}_n$n.err
done
run --config 3 --mode split --steps 10 --equal-rows > $out/${tag}_bench_split_equal_config3_n$n.json 2>> $out/${tag}_split_config3_n$n.err
run > $out/${tag}_bench_views_n$n.json 2> $out/${tag}_bench_views_n$n.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
  tools/split_check.py 0 3 > $out/${tag}_split_check_n$n.log 2>&1
tail -n 3 $out/${tag}_split_check_n$n.log
tail -c 400 $out/${tag}_*.err
