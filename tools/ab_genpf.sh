#!/bin/bash
# A/B of the generated kernels' sweep chunking (gen_l2_mb) and L2 prefetch distance (gen_pf) on one B200.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
YB_GEN_PF=1 YB_GEN_L2_MB=2 timeout 600 python -m pytest tests/test_generated_gpu.py -m gpu -x -q 2>&1 | tail -3
for l2 in 0 8 16 24 40; do for pf in 0 1; do
  timeout 120 python bench_stencils.py 512 gen_pf=$pf gen_l2_mb=$l2 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('l2=$l2 pf=$pf', d['stencil'], d['gpoints_per_s'], d['roofline_frac_of_measured_hbm'])"
done; done | tee gpurun_out/ab_genpf.txt
