#!/bin/bash
# bench with frames in flight at every N (programmatic dependent launch off with several handles): N=1 here
tag=${1:-r3d}
out=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_config3.json 2> $out/${tag}_bench.err
echo "exit $?"; tail -3 $out/${tag}_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r3d_bench_config3.json'))
print(d['value'], d['ms_per_step'], d.get('frames_in_flight'), d.get('one_frame_at_a_time'), 'e2e', d['e2e'], 'sustained', d['sustained'])
print(d['config']); print(d['roofline']); print(d['cpu_baseline'])
PY
timeout 600 python -m pytest tests -m gpu -x -q -k "smoke or facade or render_options or instance_culling" 2>&1 | tail -2
