#!/bin/bash
# r4d: the whole GPU suite (incl. comparators, clustered scene, the SimpleRenderer facade) and smoke()
tag=${1:-r4d}
timeout 400 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -15 > gpurun_out/${tag}_pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/${tag}_smoke.log
cat gpurun_out/${tag}_pytest_gpu.log gpurun_out/${tag}_smoke.log
