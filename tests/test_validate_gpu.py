"""`-validate` of the command-line harness for emitter-generated solutions: the TMA sweep kernels against the one-thread-per-point
direct kernels on identical, non-constant data, every written var compared element for element -- the role the reference's scalar
run_ref() plays behind its own `-validate` (/root/reference/src/kernel/yask_main.cpp:562-644, lib/context.cpp:85-217)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exe(stencil):
    return os.path.join(ROOT, "yask_b200", "bin", f"yask_kernel.{stencil}.b200.exe")


@pytest.mark.parametrize("stencil,args", [("awp_elastic", ["-g", "64"]), ("ssg", ["-gx", "48", "-gy", "40", "-gz", "64"]), ("3axis", ["-g", "64"])])
def test_harness_validate_generated_solution(stencil, args):
    if not os.path.exists(_exe(stencil)):
        pytest.skip("harness not built")
    r = subprocess.run([_exe(stencil)] + args + ["-validate"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "TEST PASSED" in r.stdout and "YASK DONE" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if "TEST PASSED" in l][0]
    assert int(line.split("mismatch(es) in")[1].split()[0]) > 0          # something was compared


@pytest.mark.parametrize("radius", [1, 2])
def test_harness_temporal_tile_through_the_reference_command_line(radius):
    """`-bt 2` on a radius-suffixed iso3dfd library (the reference fixes the radius at build time): -validate compares the
    temporal tile with the one-thread-per-point kernel of a second solution; a timing run reports the block steps."""
    exe = _exe(f"iso3dfd_r{radius}")
    if not os.path.exists(exe):
        pytest.skip("harness not built")
    r = subprocess.run([exe, "-g", "96", "-bt", "2", "-validate"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "TEST PASSED: 0 mismatch" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([exe, "-g", "128", "-bt", "2", "-trial_steps", "6", "-num_trials", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "YASK DONE" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert [l for l in r.stdout.splitlines() if "num-temporal-block-steps" in l][0].split()[-1] == "2"
