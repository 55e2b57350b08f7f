#!/bin/bash
# N = 2 (gpurun --gpus 2): owned-bins read-back test, split bench with frames in flight and the host-gathered e2e
tag=${1:-r2s}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "owned_bins or instance_culling or bin_range" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log; tail -3 $out/${tag}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 5 --no-views > $out/${tag}_bench_split_n2.json 2> $out/${tag}_bench_split_n2.err
echo "exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_bench_split_n2.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['one_frame_at_a_time']['value'], d['stage_ms'])
PY
tail -5 $out/${tag}_bench_split_n2.err
