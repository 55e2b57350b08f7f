#!/bin/bash
# programmatic dependent launch on / off: N=1 bench line of configs 3 and 1, stage probe
tag=${1:-r3b}
out=gpurun_out
for c in 3 1; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_config${c}_pdl.json 2> $out/${tag}_bench.err
  LUCID_NO_PDL=1 timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_config${c}_nopdl.json 2>> $out/${tag}_bench.err
done
python - <<'PY'
import json
for c in (3, 1):
    for m in ("pdl", "nopdl"):
        d = json.load(open('gpurun_out/r3b_bench_config%d_%s.json' % (c, m)))
        print(c, m, d['value'], d['ms_per_step'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], 'frame', d['stage_ms']['frame'], 'staged', d['stage_ms']['frame_with_stage_events'])
PY
tail -3 $out/${tag}_bench.err
