#!/bin/bash
# round 2, GPU call G (4 GPUs): copy-engine halo transfer -- multi-rank tests, then N=1 and N=4 with it on (default) / off
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_multi_gpu.py tests/test_cpp_api.py tests/test_iso3dfd_gpu.py -m gpu -q -k "rank or physical or two_processes or yask_sh or in_run or var_checks or fuse" --timeout=300 --timeout-method=thread > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -5 gpurun_out/g_pytest.log
B="--steps 40 --warmup 5 --no-cpu --no-e2e --no-secondary"
timeout 300 python bench.py $B > gpurun_out/g_n1.json 2> gpurun_out/g_n1.err
i=0
for opt in "" "--opt dma_halo=0" "" "--opt dma_halo=0"; do
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29630+i)) bench.py --gpus 4 $B $opt > gpurun_out/g_n4_$i.json 2> gpurun_out/g_n4_$i.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/g_n*.json")):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], (l.get("sustained") or {}).get("ms_per_step"), l.get("halo_check"), l.get("per_rank_ms_per_step"))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
P
