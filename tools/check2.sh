#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for t in "awp_elastic-n0-3-16" "cube-n5-2-16"; do
  r=$(timeout 60 python -m pytest "tests/test_generated_gpu.py::test_sweep_variant_vs_oracle[$t]" -m gpu -q --timeout 50 -p no:cacheprovider 2>&1 | tail -1)
  echo "$t: $r"
done
timeout 300 python -m pytest tests/test_generated_gpu.py -m gpu -k "sweep_variant" -q --timeout 100 -p no:cacheprovider > $O/final_pytest_sweep.log 2>&1; grep -E "^(FAILED|ERROR)" $O/final_pytest_sweep.log | cut -c1-160 | head; tail -1 $O/final_pytest_sweep.log; grep -m3 "E  " $O/final_pytest_sweep.log | cut -c1-250
for lx in 64 128 512; do timeout 100 python bench_stencils.py 512 gen_sweep=1 gen_sweep_lx=$lx 2>&1 | grep "^{" | tee -a $O/final_bench_stencils_sweep.json | cut -c1-170; done
