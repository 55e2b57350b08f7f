#!/bin/bash
# r4e: comparator tests with the final library, then two of them (order recovery on mixed instances; lists over 1024
# entries in the global sort scratch) under compute-sanitizer memcheck
tag=${1:-r4e}
timeout 100 python -m pytest tests/test_comparators.py tests/test_clustering.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest.log
timeout 110 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_comparators.py -m gpu -x -q -k "mixed_order or long_lists" > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${tag}_memcheck.log
cat gpurun_out/${tag}_pytest.log; tail -12 gpurun_out/${tag}_memcheck.log
