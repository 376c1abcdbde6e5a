#!/bin/bash
# round 2, GPU call J (1 GPU): final validation of the tree -- full GPU suite, smoke(), bench, launch list and --set full capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -u -m pytest tests -m gpu -q --maxfail=30 --timeout=300 --timeout-method=thread > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
tail -6 gpurun_out/j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/j_smoke.log; cat gpurun_out/j_smoke.log
timeout 600 python bench.py --no-cpu > gpurun_out/j_bench_n1.json 2> gpurun_out/j_bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-sustained > gpurun_out/j_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:iso3dfd_tma2 -s 2 -c 1 -o gpurun_out/j_iso_full python tools/prof_iso.py 1024 4 kernel=tma > gpurun_out/j_ncu_iso.log 2>&1
python - <<'P'
import json
l=json.loads(open("gpurun_out/j_bench_n1.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["roofline"]["frac"], l["sustained"]["value"], l["e2e"]["value"], [(s.get("value"), (s.get("roofline") or {}).get("frac")) for s in l["secondary"]])
P
