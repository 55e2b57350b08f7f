set -x
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_gpu.log
LUCID_SHADE_STREAM=ldg timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or render_options or full_size_config" > gpurun_out/r2c_pytest_gpu_ldg.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_gpu_ldg.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_config3_tma.json 2> gpurun_out/r2c_bench_config3_tma.err
LUCID_SHADE_STREAM=ldg timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_config3_ldg.json 2> gpurun_out/r2c_bench_config3_ldg.err
timeout 600 python tools/gpu_probe.py 0 1 2 3 > gpurun_out/r2c_probe_tma.txt 2>&1
LUCID_SHADE_STREAM=ldg timeout 600 python tools/gpu_probe.py 1 2 3 > gpurun_out/r2c_probe_ldg.txt 2>&1
