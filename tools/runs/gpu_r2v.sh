#!/bin/bash
# debug-raster variant, compact block lists: tests; stage times of configs 1-3 with fixed and compact lists; free memory
tag=${1:-r2v}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "debug_raster or compact" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log; tail -4 $out/${tag}_pytest.log
timeout 600 python tools/gpu_probe.py 1 2 3 > $out/${tag}_probe_fixed.txt 2>&1
timeout 600 python tools/gpu_probe.py 1 2 3 --compact > $out/${tag}_probe_compact.txt 2>&1
grep -E "^==|stage_ms" $out/${tag}_probe_fixed.txt $out/${tag}_probe_compact.txt
python - <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from lucid_b200 import api
for name, kw in (("fixed", {}), ("compact 2^25", dict(create_flags=api.CREATE_COMPACT_LISTS, max_block_entries=1 << 25))):
    torch.cuda.synchronize(); f0, _ = torch.cuda.mem_get_info()
    r = api.LucidRenderer(3840, 2160, 0, 0, **kw)
    f1, _ = torch.cuda.mem_get_info()
    print("handle memory 3840x2160", name, round((f0 - f1) / 2**30, 2), "GiB")
    r.close()
PY
