#!/bin/bash
# round 2, GPU call M (1 GPU): x neighbours in registers (emitter "x queue") -- generated-kernel suite, then an A/B on one box
# of ssg 512^3 against a library built with YB_EMIT_SWEEP_XQ=0 (tools/ab/, not committed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_generated_gpu.py -m gpu -q --maxfail=30 --timeout=300 --timeout-method=thread > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -8 gpurun_out/m_pytest.log
: > gpurun_out/m_ab.log
for rep in 1 2 3; do
  for v in xq noxq; do
    if [ $v = noxq ]; then export YASK_B200_LIB=$PWD/tools/ab/libyask_b200_noxq.so; else unset YASK_B200_LIB; fi
    echo -n "$v " >> gpurun_out/m_ab.log
    timeout 200 python -c "
import sys; sys.path.insert(0, '.')
from bench_stencils import run
r = run('ssg', 512, 30, 5, 2)
print(r['gpoints_per_s'], r['ms_per_step'], r['roofline_frac_of_measured_hbm'])" >> gpurun_out/m_ab.log 2>&1
  done
done
unset YASK_B200_LIB
cat gpurun_out/m_ab.log
timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err
python - <<'P'
import json
l=json.loads(open("gpurun_out/m_bench_n1.json").read().strip().splitlines()[-1])
print(l["value"], l["roofline"]["frac"], [(s.get("value"), (s.get("roofline") or {}).get("frac")) for s in l["secondary"]])
P
