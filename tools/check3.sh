#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for t in "cube-n5-2-16" "test_stages_3d-n6-3-8" "test_stream_3d-n8-3-8" "test_partial_3d-n7-2-16" "tti-n3-2-16" "awp-n2-2-8" "awp_elastic-n1-2-128"; do
  r=$(timeout 60 python -m pytest "tests/test_generated_gpu.py::test_sweep_variant_vs_oracle[$t]" -m gpu -q --timeout 50 -p no:cacheprovider 2>&1 | tail -1)
  echo "$t: $r"
done
cat > /tmp/sw_dbg.py <<'PY'
import sys
sys.path.insert(0, '.')
from tests.test_generated_gpu import synth_inputs, run_gpu
ins, ir = synth_inputs("awp_elastic", (12, 9, 140), 33)
out, _ = run_gpu("awp_elastic", (12, 9, 140), 1, ins, 0, opts=(("gen_sweep", 1), ("gen_sweep_lx", 8)))
print("ran ok")
PY
timeout 200 compute-sanitizer --tool memcheck --print-limit 3 python /tmp/sw_dbg.py > $O/sw_sanitizer.log 2>&1; grep -v "^=========     Host Frame\|^=========         in " $O/sw_sanitizer.log | head -40
