#!/bin/bash
# round 2, GPU call P (1 GPU): ncu --set full of the ssg sweep kernels with the x queue (refresh of profiles/r2_ssg_sweep_ncu.txt)
# and the same-box bench_stencils line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python bench_stencils.py 512 gen_sweep=1 > gpurun_out/p_bench_stencils.json 2> gpurun_out/p_bench_stencils.err
cat gpurun_out/p_bench_stencils.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ssg_part -s 2 -c 2 -o gpurun_out/p_ssg_sweep python tools/prof_gen.py ssg 512 gen_sweep=1 > gpurun_out/p_ncu_ssg.log 2>&1
tail -3 gpurun_out/p_ncu_ssg.log
python tools/ncu_summary.py gpurun_out/p_ssg_sweep.ncu-rep > gpurun_out/p_ssg_sweep_ncu.txt 2>&1
head -60 gpurun_out/p_ssg_sweep_ncu.txt
