#!/bin/bash
# after the shared-memory tie order: whole GPU suite, smoke, stage times, split played on one GPU, N=1 bench
tag=${1:-r2x}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log; tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout 600 python tools/gpu_probe.py 1 2 3 > $out/${tag}_probe_stage_times.txt 2>&1
grep -E "^==|stage_ms" $out/${tag}_probe_stage_times.txt | cut -c1-330
timeout 600 python tools/split_probe.py 3 8 --cull > $out/${tag}_split_probe.txt 2>&1
cut -c1-330 $out/${tag}_split_probe.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_config3.json 2> $out/${tag}_bench.err
python -c "
import json; d=json.load(open('$out/${tag}_bench_config3.json')); print(d['value'], d['stage_ms'], d['e2e']['value'], d['sustained']['value'], d.get('cpu_baseline'))"
