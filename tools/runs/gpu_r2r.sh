#!/bin/bash
# frames in flight in the split bench, N = 2 (gpurun --gpus 2)
tag=${1:-r2r}
out=gpurun_out
mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 5 --no-views > $out/${tag}_bench_split_n2.json 2> $out/${tag}_bench_split_n2.err
echo "exit $?"; cat $out/${tag}_bench_split_n2.json | cut -c1-3000; tail -5 $out/${tag}_bench_split_n2.err
