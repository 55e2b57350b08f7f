#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): split bench at N = 8 (with the views line) and N = 4
out=gpurun_out
tag=${1:-r2z}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --trace-split > $out/${tag}_bench_split_n8.json 2> $out/${tag}_bench_split_n8.err
echo "n8 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 5 --no-views > $out/${tag}_bench_split_n4.json 2> $out/${tag}_bench_split_n4.err
echo "n4 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --no-views > $out/${tag}_bench_split_n2.json 2> $out/${tag}_bench_split_n2.err
echo "n2 exit $?"
python - $tag <<'PY'
import json, sys
for n in (8, 4, 2):
    try:
        d = json.load(open('gpurun_out/%s_bench_split_n%d.json' % (sys.argv[1], n)))
        print(n, d['value'], d['ms_per_step'], 'serial', d['one_frame_at_a_time']['value'], 'e2e', d['e2e']['value'], d['e2e'].get('gathered_on_rank0'), d['e2e'].get('verified_against_gathered_frame'))
        print('  ', d['stage_ms'], d['rank_frame_ms'])
        if 'views' in d: print('   views', d['views']['value'], d['views']['e2e'])
    except Exception as e:
        print(n, 'failed', e)
PY
tail -4 $out/${tag}_bench_split_n8.err
