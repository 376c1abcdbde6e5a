#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
YB_GEN_SWEEP=1 timeout 150 python -m pytest tests/test_generated_gpu.py tests/test_cpp_api.py tests/test_multi_gpu.py -m gpu -q --maxfail=10 --timeout 60 -p no:cacheprovider > $O/pytest_sweep_default.log 2>&1; echo "rc=$?"; grep -E "^(FAILED|ERROR)" $O/pytest_sweep_default.log | cut -c1-180 | head; tail -1 $O/pytest_sweep_default.log
