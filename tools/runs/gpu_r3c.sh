#!/bin/bash
# programmatic dependent launch on / off / auto on the ranks of an 8-way split played on one GPU (1..5 frames in flight),
# and the N=1 line of configs 3, 2, 1 with the automatic choice
tag=${1:-r3c}
out=gpurun_out
for m in on off auto; do
  LUCID_PDL=$m PROBE_HANDLES=5 timeout 900 python tools/inflight_probe.py 3 40 > $out/${tag}_inflight_pdl_$m.txt 2>&1
  echo "PDL $m"; tail -4 $out/${tag}_inflight_pdl_$m.txt | cut -c1-420
done
for c in 3 2 1 0; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_config${c}.json 2> $out/${tag}_bench.err
done
python - <<'PY'
import json
for c in (3, 2, 1, 0):
    d = json.load(open('gpurun_out/r3c_bench_config%d.json' % c))
    print(c, d['value'], d['ms_per_step'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], 'frame', d['stage_ms']['frame'], 'staged', d['stage_ms']['frame_with_stage_events'])
PY
