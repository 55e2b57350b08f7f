# 8-GPU measurement pass (gpurun --gpus 8): split bench at N = 8 / 4 / 2 and the host read-back ceiling
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --trace-split > gpurun_out/r2k_bench_split_n8.json 2> gpurun_out/r2k_bench_split_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/d2h_ceiling.py 1920 1080 > gpurun_out/r2k_d2h_ceiling_n8.json 2> gpurun_out/r2k_d2h_ceiling.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29514 tools/d2h_ceiling.py 1920 1080 > gpurun_out/r2k_d2h_ceiling_n1.json 2>> gpurun_out/r2k_d2h_ceiling.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 tools/d2h_ceiling.py 1920 1080 > gpurun_out/r2k_d2h_ceiling_n4.json 2>> gpurun_out/r2k_d2h_ceiling.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 5 --no-views > gpurun_out/r2k_bench_split_n2.json 2> gpurun_out/r2k_bench_split_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 5 --no-views > gpurun_out/r2k_bench_split_n4.json 2> gpurun_out/r2k_bench_split_n4.err
