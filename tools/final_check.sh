#!/bin/bash
# One-box verification of the round's final state: GPU tests, smoke, both bench arms, secondary stencils, one ncu capture
# of the generated awp_elastic kernels.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests/test_generated_gpu.py tests/test_cpp_api.py tests/test_iso3dfd_gpu.py tests/test_multi_gpu.py tests -m gpu -k "not sweep_variant" -q --maxfail=12 --timeout 150 -p no:cacheprovider > $O/final_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" $O/final_pytest.log | cut -c1-200 | head -14; tail -2 $O/final_pytest.log
timeout 200 python -m pytest tests/test_generated_gpu.py -m gpu -k sweep_variant -q --timeout 100 -p no:cacheprovider > $O/final_pytest_sweep.log 2>&1; echo "sweep pytest rc=$?"; grep -E "^(FAILED|ERROR)" $O/final_pytest_sweep.log | cut -c1-160 | head; tail -1 $O/final_pytest_sweep.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>$O/final_bench.err | grep "^{" > $O/final_bench_n1.json; cut -c1-200 $O/final_bench_n1.json
timeout 300 python bench_stencils.py 512 2>&1 | grep "^{" > $O/final_bench_stencils.json; cut -c1-170 $O/final_bench_stencils.json
timeout 200 python bench_stencils.py 512 gen_sweep=1 2>&1 | grep "^{" > $O/final_bench_stencils_sweep.json; cut -c1-170 $O/final_bench_stencils_sweep.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:awp_elastic_part -c 2 -o $O/awp_r1final -f python tools/prof_gen.py awp_elastic 512 > $O/final_ncu.log 2>&1; echo "ncu rc=$?"
timeout 120 yask_b200/bin/yask_kernel.iso3dfd.b200.exe -g 1024 -trial_steps 50 -num_trials 3 -no-pre_auto_tune > $O/final_harness_iso3dfd.log 2>&1; grep "best-throughput (num-points" $O/final_harness_iso3dfd.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/final_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
