#!/bin/bash
# round 2, GPU call H (4 GPUs): interleaved A/B of the halo transfer at N=4 -- copy engines (default) / the kernel's own peer
# stores (dma_halo=0) / no transfer at all (peer_probe=1, timing floor) -- burst (40 steps) and sustained (>= 2 s) values
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B="--steps 40 --warmup 5 --no-cpu --no-e2e --no-secondary --no-halo-check"
timeout 300 python bench.py $B > gpurun_out/h_n1.json 2> gpurun_out/h_n1.err
i=0
for rep in 1 2; do
for opt in "--opt dma_halo=1" "--opt dma_halo=0" "--opt dma_halo=0 --opt peer_probe=1"; do
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29640+i)) bench.py --gpus 4 $B $opt > gpurun_out/h_n4_$i.json 2> gpurun_out/h_n4_$i.err
done
done
timeout 300 python bench.py $B > gpurun_out/h_n1b.json 2> gpurun_out/h_n1b.err
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/h_n*.json")):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], (l.get("sustained") or {}).get("ms_per_step"), (l.get("sustained") or {}).get("clocks",{}).get("sm_mhz"), l.get("per_rank_ms_per_step"))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
P
