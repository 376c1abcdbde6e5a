"""Python module `yask_kernel` (pybind11 over the C++ mirror, yask_b200/csrc/yk_pybind.cpp) -- what can be checked without a GPU:
the module of every built solution imports under the reference's name, exposes the reference's classes
(/root/reference/src/kernel/swig/yask_kernel_api.i), maps yask_exception to RuntimeError and refuses to prepare without a device."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODS = sorted(glob.glob(os.path.join(ROOT, "yask_b200", "lib", "python", "*", "yask_kernel*.so")))

SCRIPT = r'''
import yask_kernel as yk
kfac = yk.yk_factory(); ofac = yk.yask_output_factory()
env = kfac.new_env()
assert env.get_num_ranks() == 1 and env.get_rank_index() == 0
soln = kfac.new_solution(env)
env.set_debug_output(ofac.new_string_output())
dims = soln.get_domain_dim_names()
for d in dims:
    soln.set_overall_domain_size(d, 64); soln.set_min_pad_size(d, 1); soln.set_block_size(d, 32)
assert soln.get_overall_domain_size_vec() == [64] * len(dims)
fv = soln.new_fixed_size_var("fvar", dims, [5] * len(dims))
assert fv.is_fixed_size() and fv.get_name() == "fvar"
assert yk.cvar.yask_numa_local == -1 and yk.cvar.yask_numa_none == -9
n = 0
for call in (lambda: soln.run_solution(0), lambda: soln.run_auto_tuner_now(False), lambda: soln.prepare_solution()):
    try:
        call()
    except RuntimeError as e:
        n += 1
        assert "YASK error" in str(e)
assert n == 3
assert soln.apply_command_line_options("-bt 2 -unknown 3") == "-unknown 3"
assert all(v.get_dim_names() for v in soln.get_vars() if not v.is_fixed_size()) or soln.get_num_vars() >= 1
print("OK", soln.get_name(), soln.get_element_bytes(), soln.get_step_dim_name(), [v.get_name() for v in soln.get_vars()])
'''


@pytest.mark.skipif(not MODS, reason="python modules not built (pybind11 missing?)")
@pytest.mark.parametrize("mod", MODS, ids=[os.path.basename(os.path.dirname(m)) for m in MODS])
def test_module_mirrors_reference_python_api(mod):
    env = dict(os.environ, PYTHONPATH=os.path.dirname(mod), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    stencil = os.path.basename(os.path.dirname(mod))
    if stencil.startswith("iso3dfd_r"):          # radius-suffixed library of iso3dfd (the reference fixes the radius at build time)
        stencil = "iso3dfd"
    assert r.stdout.strip().startswith("OK " + stencil)
