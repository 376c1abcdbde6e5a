#!/bin/bash
# round 2, GPU call K (8 GPUs): the 8-GPU bench line (halo check, secondary workloads) and the independent-replicas diagnostic
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/k_bench_n8.json 2> gpurun_out/k_bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 8 --steps 30 --warmup 5 --replicas --no-e2e --no-secondary > gpurun_out/k_bench_n8_replicas.json 2> gpurun_out/k_bench_n8_replicas.err
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-secondary > gpurun_out/k_bench_n1.json 2> gpurun_out/k_bench_n1.err
python - <<'P'
import json
for f in ("k_bench_n8","k_bench_n8_replicas","k_bench_n1"):
    try:
        l=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], (l.get("sustained") or {}).get("ms_per_step"), l.get("halo_check"), l.get("per_rank_ms_per_step"), (l.get("e2e") or {}).get("value"), [(s.get("value"), (s.get("roofline") or {}).get("frac")) for s in (l.get("secondary") or [])])
    except Exception as e: print(f, "ERR", e, open("gpurun_out/%s.err"%f).read()[-300:])
P
