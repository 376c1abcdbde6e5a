"""The reference's own Python API tests, UNMODIFIED (/root/reference/src/kernel/tests/yask_kernel_api_test.py and
yask_kernel_api_exception_test.py, staged by the build under tests/_bin like the compiled C++ API tests), run against the
`yask_kernel` module of this engine (pybind11, yask_b200/csrc/yk_pybind.cpp) on the GPU: vars are filled through NumPy
buffers, read back, the solution runs 1 + 4 steps, stats are read -- every assert in those scripts must hold."""
import glob
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "_bin")
MODS = sorted(glob.glob(os.path.join(ROOT, "yask_b200", "lib", "python", "*", "yask_kernel*.so")))


def _run(script, mod):
    path = os.path.join(BIN, script)
    if not os.path.exists(path):
        pytest.skip(f"{script} was not staged (built outside the container that holds the reference)")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(mod))
    return subprocess.run([sys.executable, path], env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(not MODS, reason="python modules not built")
@pytest.mark.parametrize("mod", MODS, ids=[os.path.basename(os.path.dirname(m)) for m in MODS])
def test_reference_python_api_test_unmodified(mod):
    r = _run("ref_yask_kernel_api_test.ref.py", mod)
    assert r.returncode == 0, (r.stdout[-3000:] + r.stderr[-3000:])
    assert "End of YASK Python kernel API test." in r.stdout


@pytest.mark.skipif(not MODS, reason="python modules not built")
def test_reference_python_exception_test_unmodified():
    mod = [m for m in MODS if os.sep + "iso3dfd" + os.sep in m] or MODS
    r = _run("ref_yask_kernel_api_exception_test.ref.py", mod[0])
    assert r.returncode == 0, (r.stdout[-3000:] + r.stderr[-3000:])
    assert "End of YASK Python kernel API test with exception." in r.stdout
