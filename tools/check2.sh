#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_generated_gpu.py -m gpu -k "sweep_variant" -q --timeout 100 -p no:cacheprovider > $O/final_pytest_sweep.log 2>&1; grep -E "^(FAILED|ERROR)" $O/final_pytest_sweep.log | cut -c1-160 | head; tail -1 $O/final_pytest_sweep.log; grep -m3 "E  " $O/final_pytest_sweep.log | cut -c1-250
rm -f $O/final_bench_stencils_sweep.json
timeout 100 python bench_stencils.py 512 gen_sweep=0 2>&1 | grep "^{" | tee -a $O/final_bench_stencils_sweep.json | cut -c1-170
for lx in 128 256; do timeout 100 python bench_stencils.py 512 gen_sweep=1 gen_sweep_lx=$lx 2>&1 | grep "^{" | tee -a $O/final_bench_stencils_sweep.json | cut -c1-170; done
cat > /tmp/prof_sw.py <<'PY'
import sys
sys.path.insert(0, '.')
from bench_stencils import run
print(run("awp_elastic", 512, 3, 1, 2, ["gen_sweep=1"]))
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:awp_elastic_part --csv --log-file $O/sweep_launches.csv python /tmp/prof_sw.py > /dev/null 2>&1; grep "awp_elastic_part" $O/sweep_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -4
