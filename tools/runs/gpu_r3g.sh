#!/bin/bash
tag=${1:-r3g}
out=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "frames_in_flight or owned_bins" 2>&1 | tail -2
timeout 600 python bench.py --config 2 --steps 10 --no-cpu-baseline > $out/${tag}_bench_config2.json 2> $out/${tag}_bench.err; echo "exit $?"
timeout 600 python bench.py > $out/${tag}_bench_config3.json 2>> $out/${tag}_bench.err; echo "exit $?"
python - <<'PY'
import json
for c in (2, 3):
    d = json.load(open('gpurun_out/r3g_bench_config%d.json' % c))
    print(c, d['value'], d['ms_per_step'], d.get('frames_in_flight'), d.get('frames_in_flight_tried'), 'e2e', d['e2e'], 'roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['sm_issue_active_pct'], d['gpu_launches'])
PY
tail -2 $out/${tag}_bench.err
