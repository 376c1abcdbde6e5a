#!/bin/bash
port=29600
for rep in 1 2; do for f in 1 0; do
port=$((port+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 40 --warmup 3 --no-e2e --opt fused_halo=$f 2>gpurun_out/ab_err_$rep$f.log | grep "^{" | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('fused=$f', d['value'], d['ms_per_step'], d['gpu_launches'])" || tail -3 gpurun_out/ab_err_$rep$f.log
done; done
