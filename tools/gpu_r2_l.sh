#!/bin/bash
# round 2, GPU call L (1 GPU): validation of the tree after call J (re-emitted sweep kernels with byte-offset ring positions,
# wave2d contraction order, fp64 iso3dfd radius 3 / sponge radius 6, step-wrapped reads) -- full GPU suite, smoke(), bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu -q --maxfail=30 --timeout=300 --timeout-method=thread > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
tail -8 gpurun_out/l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/l_smoke.log; cat gpurun_out/l_smoke.log
timeout 600 python bench.py --no-cpu > gpurun_out/l_bench_n1.json 2> gpurun_out/l_bench_n1.err
python - <<'P'
import json
l=json.loads(open("gpurun_out/l_bench_n1.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["roofline"]["frac"], l["sustained"]["value"], l["e2e"]["value"], [(s.get("value"), (s.get("roofline") or {}).get("frac")) for s in l["secondary"]])
P
