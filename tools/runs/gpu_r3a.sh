#!/bin/bash
# 8 GPUs: split bench with the in-flight rebalancing, 5 (default, with the views line) and 7 frames in flight
out=gpurun_out
tag=${1:-r3a}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > $out/${tag}_bench_split_n8.json 2> $out/${tag}_bench_split_n8.err
echo "n8 exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-views --frames-in-flight 7 > $out/${tag}_bench_split_n8_f7.json 2> $out/${tag}_bench_split_n8_f7.err
echo "n8 f7 exit $?"
python - <<'PY'
import json
for f in ("r3a_bench_split_n8", "r3a_bench_split_n8_f7"):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, d['value'], d['ms_per_step'], 'serial', d['one_frame_at_a_time']['value'], 'e2e', d['e2e']['value'], d['sustained']['value'], d['config']['parallelism'][-60:])
    except Exception as e:
        print(f, 'failed', e)
PY
tail -3 $out/${tag}_bench_split_n8.err
