#!/bin/bash
# round 2, GPU call F (4 GPUs): what does the multi-rank step cost?  N=1, then N=4 normal / peer stores skipped (timing probe) /
# skewed CTA starts / round-1 order (overlap off)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B="--steps 40 --warmup 5 --no-cpu --no-e2e --no-secondary --no-sustained"
timeout 300 python bench.py $B > gpurun_out/f_n1.json 2> gpurun_out/f_n1.err
i=0
for opt in "" "--opt peer_probe=1 --no-halo-check" "--opt peer_probe=2" "--opt overlap_comms=0" ""; do
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29620+i)) bench.py --gpus 4 $B $opt > gpurun_out/f_n4_$i.json 2> gpurun_out/f_n4_$i.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/f_n*.json")):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], l.get("halo_check"), l.get("per_rank_ms_per_step"))
    except Exception as e: print(f, "ERR", e)
P
