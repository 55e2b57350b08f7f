#!/bin/bash
# Multi-GPU pass (run under gpurun --gpus N): views-sharded bench (the driver's scaling run) and the
# bin-row split of one 4K frame with P2P stores into rank 0's image.   bash tools/gpu_multi.sh <tag> <N>
tag=${1:-r1}
n=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
timeout 300 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
run > $out/${tag}_bench_views_n$n.json 2> $out/${tag}_bench_views_n$n.err
run --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_n$n.json 2> $out/${tag}_bench_reference_n$n.err
for c in 3 2; do
  timeout 300 python bench.py --config $c --steps 10 --no-cpu-baseline > $out/${tag}_bench_config${c}_n1.json 2> $out/${tag}_bench_config${c}_n1.err
  run --config $c --mode split --steps 10 > $out/${tag}_bench_split_config${c}_n$n.json 2> $out/${tag}_bench_split_config${c}_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
  tools/split_check.py > $out/${tag}_split_check_n$n.log 2>&1
tail -c 600 $out/${tag}_*.err
