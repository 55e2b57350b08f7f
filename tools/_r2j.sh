set -x
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest_gpu.log
