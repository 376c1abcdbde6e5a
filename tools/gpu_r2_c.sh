#!/bin/bash
# round 2, GPU call C (1 GPU): sweep kernels forced on across the generated suite (per-test timeout), new API tests, ssg A/B,
# iso3dfd L2-policy experiment, full N=1 bench with the CPU baselines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
YB_GEN_SWEEP=1 timeout 900 python -u -m pytest tests/test_generated_gpu.py -m gpu -v -x --timeout=150 --timeout-method=thread > gpurun_out/c_pytest_sweep.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest_sweep.log
tail -4 gpurun_out/c_pytest_sweep.log
timeout 900 python -u -m pytest tests/test_cpp_api.py tests/test_iso3dfd_gpu.py tests/test_multi_gpu.py -m gpu -q --timeout=200 --timeout-method=thread > gpurun_out/c_pytest_rest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest_rest.log
tail -4 gpurun_out/c_pytest_rest.log
for sw in 0 1; do
  timeout 300 python bench_stencils.py 512 gen_sweep=$sw >> gpurun_out/c_bench_stencils.json 2>> gpurun_out/c_bench_stencils.err
done
cat gpurun_out/c_bench_stencils.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ssg_part -s 2 -c 2 -o gpurun_out/c_ssg_sweep python tools/prof_gen.py ssg 512 gen_sweep=1 > gpurun_out/c_ncu_ssg.log 2>&1
for cfg in "st_cs=1" "pol_h=2" "pol_c=2" "st_cs=1 pol_c=2 pol_h=2"; do
  tag=$(echo $cfg | tr ' =' '__')
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:iso3dfd_tma2 -s 1 -c 2 --csv --log-file gpurun_out/c_iso_$tag.csv python tools/prof_iso.py 1024 3 kernel=tma $cfg > gpurun_out/c_iso_$tag.log 2>&1
done
timeout 700 python bench.py > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err
tail -c 1500 gpurun_out/c_bench_n1.json
