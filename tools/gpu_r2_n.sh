#!/bin/bash
# round 2, GPU call N (2 GPUs): multi-rank tests and the two-GPU bench line with the x-queue sweep kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 500 python -u -m pytest tests/test_multi_gpu.py tests/test_cpp_api.py -m gpu -q --maxfail=10 --timeout=300 --timeout-method=thread > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -6 gpurun_out/n_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu > gpurun_out/n_bench_n2.json 2> gpurun_out/n_bench_n2.err
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-secondary > gpurun_out/n_bench_n1.json 2> gpurun_out/n_bench_n1.err
python - <<'P'
import json
for f in ("gpurun_out/n_bench_n2.json", "gpurun_out/n_bench_n1.json"):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], l.get("halo_check"), l.get("per_rank_ms_per_step"), [(s.get("value"), (s.get("roofline") or {}).get("frac"), s.get("error")) for s in (l.get("secondary") or [])])
    except Exception as e: print(f, "ERR", e)
P
