#!/bin/bash
# round 2, GPU call D (2 GPUs): full GPU suite on the current tree, iso3dfd L2-policy A/B on one box, bench at N=1 (no CPU) and N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -u -m pytest tests -m gpu -q --maxfail=30 --timeout=300 --timeout-method=thread > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -6 gpurun_out/d_pytest.log
for cfg in "" "st_cs=1 pol_c=2 pol_h=2" "st_cs=1" "" "st_cs=1 pol_c=2 pol_h=2"; do
  echo "cfg: $cfg" >> gpurun_out/d_iso_ab.log
  timeout 120 python tools/prof_iso.py 1024 40 kernel=tma $cfg >> gpurun_out/d_iso_ab.log 2>&1
done
cat gpurun_out/d_iso_ab.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 30 --warmup 5 --no-secondary --no-e2e --opt overlap_comms=0 > gpurun_out/d_bench_n2_nooverlap.json 2> gpurun_out/d_bench_n2_nooverlap.err
python - <<'P'
import json
for f in ("d_bench_n1","d_bench_n2","d_bench_n2_nooverlap"):
    try:
        l=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], l.get("sustained",{}) and l["sustained"].get("ms_per_step"), l.get("halo_check"), l.get("per_rank_ms_per_step"), [(s.get("value"), (s.get("roofline") or {}).get("frac")) for s in (l.get("secondary") or [])])
    except Exception as e: print(f, "ERR", e)
P
