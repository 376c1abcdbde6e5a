#!/bin/bash
# round 2, GPU call Q (1 GPU): ssg sweep kernels with 3 planes of prefetch (library in tools/ab/, not committed) against the
# default 2, same box, alternating; the ssg tests with the variant library first
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
YASK_B200_LIB=$PWD/tools/ab/libyask_b200_pf3.so timeout 300 python -u -m pytest tests/test_generated_gpu.py -m gpu -q -k ssg --timeout=200 --timeout-method=thread > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
: > gpurun_out/q_ab.log
for rep in 1 2 3; do
  for v in pf2 pf3; do
    if [ $v = pf3 ]; then export YASK_B200_LIB=$PWD/tools/ab/libyask_b200_pf3.so; else unset YASK_B200_LIB; fi
    echo -n "$v " >> gpurun_out/q_ab.log
    timeout 200 python -c "
import sys; sys.path.insert(0, '.')
from bench_stencils import run
r = run('ssg', 512, 30, 5, 2)
print(r['gpoints_per_s'], r['ms_per_step'], r['roofline_frac_of_measured_hbm'])" >> gpurun_out/q_ab.log 2>&1
  done
done
cat gpurun_out/q_ab.log
