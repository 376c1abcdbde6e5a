#!/bin/bash
# round 2, GPU call B (1 GPU): sweep kernels v2 -- correctness with the sweep forced on everywhere, A/B vs the direct kernels,
# ncu of the four benchmarked sweep kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
YB_GEN_SWEEP=1 timeout 900 python -m pytest tests/test_generated_gpu.py -m gpu -q --maxfail=30 > gpurun_out/b_pytest_sweep.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest_sweep.log
tail -5 gpurun_out/b_pytest_sweep.log
for sw in 0 1; do
  timeout 300 python bench_stencils.py 512 gen_sweep=$sw >> gpurun_out/b_bench_stencils.json 2>> gpurun_out/b_bench_stencils.err
done
cat gpurun_out/b_bench_stencils.json
for st in awp_elastic ssg; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 2 -o gpurun_out/b_${st}_sweep python tools/prof_gen.py $st 512 gen_sweep=1 > gpurun_out/b_ncu_${st}.log 2>&1
done
ls -la gpurun_out | tail -8
# iso3dfd: where does the x1.17 DRAM overfetch come from?  DRAM bytes per launch for two tile shapes and two L2-hint settings
for cfg in "tile=7" "tile=6" "tile=7 pol_c=2 pol_h=1" "tile=7 lx=1024"; do
  tag=$(echo $cfg | tr ' =' '__')
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum --clock-control none -k regex:iso3dfd_tma2 -s 1 -c 2 --csv --log-file gpurun_out/b_iso_$tag.csv python tools/prof_iso.py 1024 3 kernel=tma $cfg > gpurun_out/b_iso_$tag.log 2>&1
done
