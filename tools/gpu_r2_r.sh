#!/bin/bash
# round 2, GPU call R (1 GPU): final tree (third plane of prefetch on 4-row tiles) -- full GPU suite, smoke(), bench; then ssg with a
# fourth plane on stage 1 (library in tools/ab/, not committed) against the default, same box, alternating
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests -m gpu -q --maxfail=30 --timeout=300 --timeout-method=thread > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log
tail -5 gpurun_out/r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r_smoke.log; cat gpurun_out/r_smoke.log
: > gpurun_out/r_ab.log
for rep in 1 2; do
  for v in pf3 pf4; do
    if [ $v = pf4 ]; then export YASK_B200_LIB=$PWD/tools/ab/libyask_b200_pf4.so; else unset YASK_B200_LIB; fi
    echo -n "$v " >> gpurun_out/r_ab.log
    timeout 200 python -c "
import sys; sys.path.insert(0, '.')
from bench_stencils import run
r = run('ssg', 512, 30, 5, 2)
print(r['gpoints_per_s'], r['ms_per_step'], r['roofline_frac_of_measured_hbm'])" >> gpurun_out/r_ab.log 2>&1
  done
done
unset YASK_B200_LIB
cat gpurun_out/r_ab.log
timeout 300 python bench.py --no-cpu > gpurun_out/r_bench_n1.json 2> gpurun_out/r_bench_n1.err
python - <<'P'
import json
l=json.loads(open("gpurun_out/r_bench_n1.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["roofline"]["frac"], l["sustained"]["value"], l["e2e"]["value"], [(s.get("value"), (s.get("roofline") or {}).get("frac")) for s in l["secondary"]])
P
