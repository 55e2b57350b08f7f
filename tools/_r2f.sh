set -x
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2f_bench_config3.json 2> gpurun_out/r2f_bench_config3.err
timeout 600 python tools/gpu_probe.py 1 2 3 > gpurun_out/r2f_probe.txt 2>&1
