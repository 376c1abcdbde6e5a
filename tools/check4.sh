#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
cat > /tmp/prof_sw.py <<'PY'
import sys
sys.path.insert(0, '.')
from bench_stencils import run
print(run("awp_elastic", 512, 2, 1, 2, ["gen_sweep=1"]))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:awp_elastic_part_._sweep -c 2 -o $O/awp_sweep_r1 -f python /tmp/prof_sw.py > $O/ncu_sweep.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_sweep.log
