#!/bin/bash
# round 2, GPU call TT (1 GPU, the round's last 6 GPU-minutes): temporal tile -- bit-exactness, 1024^3 timing against the one-step
# kernels, then one ncu --set full capture of the fused kernel and of the one-step kernel at radius 2 (1024^3, one launch each)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 150 python -u tools/tt_check.py 1024 > gpurun_out/tt_check.log 2>&1; echo "tt_check rc=$?" >> gpurun_out/tt_check.log
tail -4 gpurun_out/tt_check.log
timeout 100 ncu --set full --clock-control none --import-source on -k regex:tt2_kernel -s 1 -c 1 -o gpurun_out/tt_r2 -f python tools/prof_tt.py 1024 4 2 block_steps=2 > gpurun_out/tt_ncu.log 2>&1; echo "ncu tt rc=$?" >> gpurun_out/tt_ncu.log
timeout 100 ncu --set full --clock-control none -k regex:tma2_kernel -s 1 -c 1 -o gpurun_out/tt_r2_onestep -f python tools/prof_tt.py 1024 3 2 > gpurun_out/tt_ncu1.log 2>&1; echo "ncu 1step rc=$?" >> gpurun_out/tt_ncu1.log
tail -2 gpurun_out/tt_ncu.log gpurun_out/tt_ncu1.log
ls -la gpurun_out | tail -8
