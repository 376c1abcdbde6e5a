#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_generated_gpu.py -m gpu -k "partial" -q --timeout 100 -p no:cacheprovider 2>&1 | tail -3
timeout 60 python -m pytest "tests/test_generated_gpu.py::test_sweep_variant_vs_oracle[3axis-n4-3-16]" -m gpu -q --timeout 50 -p no:cacheprovider > $O/sweep_first.log 2>&1; tail -1 $O/sweep_first.log; grep -m3 "E  " $O/sweep_first.log | cut -c1-250
timeout 300 python -m pytest tests/test_generated_gpu.py -m gpu -k "sweep_variant" -q --timeout 100 -p no:cacheprovider > $O/final_pytest_sweep.log 2>&1; grep -E "^(FAILED|ERROR)" $O/final_pytest_sweep.log | cut -c1-160 | head; tail -1 $O/final_pytest_sweep.log; grep -m3 "E  " $O/final_pytest_sweep.log | cut -c1-250
timeout 200 python bench_stencils.py 512 gen_sweep=1 2>&1 | grep "^{" > $O/final_bench_stencils_sweep.json; cut -c1-170 $O/final_bench_stencils_sweep.json
