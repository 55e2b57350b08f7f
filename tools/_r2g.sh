set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --trace-split > gpurun_out/r2g_bench_split_n8.json 2> gpurun_out/r2g_bench_split_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --completion allreduce --no-views > gpurun_out/r2g_bench_split_allreduce_n8.json 2> gpurun_out/r2g_bench_split_allreduce_n8.err
