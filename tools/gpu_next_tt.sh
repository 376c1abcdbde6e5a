#!/bin/bash
# NOT RUN YET (the round's GPU budget ended before it): the measurements the temporal tile still owes -- every form at both radii
# against the one-step sweep on one box, the offline tuner's choice, and ncu --set full of the register-queue forms at radius 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for r in 1 2; do
  timeout 200 python bench_temporal.py 1024 $r 30 > gpurun_out/next_tt_r$r.json 2> gpurun_out/next_tt_r$r.err; tail -c 600 gpurun_out/next_tt_r$r.json
done
timeout 120 python - <<'P' > gpurun_out/next_tt_tuner.log 2>&1
from yask_b200 import capi
from yask_b200.synth import var_salt
for R in (1, 2):
    s = capi.Solution("iso3dfd", radius=R); s.set_overall_domain_size_vec((1024, 1024, 1024)); s.set_option("block_steps", 2); s.prepare_solution(0)
    p, v = s.get_var("p"), s.get_var("v")
    for t in (0, 1): p.fill_hash(t, 7, var_salt("p", t), -1.0, 1.0)
    v.fill_hash(0, 7, var_salt("v", 0), 0.05, 0.3)
    print("radius", R); print(s.run_auto_tuner_now()); s.close()
P
cat gpurun_out/next_tt_tuner.log
for v in 1 3; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:tt2_kernel -s 1 -c 1 -o gpurun_out/next_tt_r2_v$v -f python tools/prof_tt.py 1024 4 2 block_steps=2 tt_variant=$v > gpurun_out/next_tt_ncu_v$v.log 2>&1
  python tools/ncu_summary.py gpurun_out/next_tt_r2_v$v.ncu-rep > gpurun_out/next_tt_r2_v${v}_ncu.txt 2>&1; head -12 gpurun_out/next_tt_r2_v${v}_ncu.txt
done
timeout 300 python -m pytest tests/test_temporal_gpu.py tests/test_python_api_gpu.py tests/test_validate_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/next_tt_pytest.log 2>&1; tail -3 gpurun_out/next_tt_pytest.log
