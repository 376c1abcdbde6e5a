#!/bin/bash
# round 2, GPU call E (4 GPUs): one-device-per-rank tests, bench at N=4 (iso3dfd + awp_elastic + ssg weak scaling: BASELINE config 5)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_multi_gpu.py tests/test_iso3dfd_gpu.py -m gpu -q -k "physical or var_checks or fuse or two_processes" --timeout=300 --timeout-method=thread > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -4 gpurun_out/e_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/e_bench_n4.json 2> gpurun_out/e_bench_n4.err
python - <<'P'
import json
for f in ("e_bench_n1","e_bench_n4"):
    try:
        l=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], l.get("sustained",{}) and l["sustained"].get("ms_per_step"), l.get("halo_check"), l.get("per_rank_ms_per_step"), (l.get("e2e") or {}).get("value"), [(s.get("value"), (s.get("roofline") or {}).get("frac")) for s in (l.get("secondary") or [])])
    except Exception as e: print(f, "ERR", e)
P
