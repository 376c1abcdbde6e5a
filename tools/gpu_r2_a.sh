#!/bin/bash
# round 2, GPU call A (2 GPUs): full GPU suite, ulp histograms, bench at N=1 and N=2 with the halo check
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python tools/ulp_hist.py > gpurun_out/a_ulp.json 2> gpurun_out/a_ulp.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/a_bench_n2.json 2> gpurun_out/a_bench_n2.err
tail -3 gpurun_out/a_pytest.log; tail -c 600 gpurun_out/a_bench_n2.json
