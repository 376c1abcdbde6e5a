"""CPU: the C-ABI library loads, exports every symbol include/yask_b200.h declares, and the host-side
logic that needs no device behaves like the reference API (errors, settings)."""
import os
import re

import pytest

from yask_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    hdr = open(os.path.join(ROOT, "include", "yask_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(capi.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_version_and_registry():
    assert capi.version().startswith("4.05.04")
    L = capi.lib()
    names = [L.yb_stencil_name(i).decode() for i in range(L.yb_num_stencils())]
    assert "iso3dfd" in names


def test_solution_settings_and_errors_without_device():
    s = capi.Solution("iso3dfd")
    assert s.get_name() == "iso3dfd" and s.get_target() == "sm_100a" and s.get_element_bytes() == 4
    assert s.get_domain_dim_names() == ["x", "y", "z"] and s.get_step_dim_name() == "t"
    s.set_overall_domain_size_vec([64, 48, 32])
    assert s.get_overall_domain_size_vec() == [64, 48, 32]
    p = s.get_var("p")
    assert p.get_dim_names() == ["t", "x", "y", "z"]
    vi = p.info
    assert vi.step_alloc == 2 and vi.dims[1].left_halo == 8 and vi.dims[3].right_halo == 8 and vi.halo_exchange_l1_norm == 1
    with pytest.raises(capi.YaskError):   # unknown var (reference: context.hpp:631-636)
        s.get_var("nope")
    with pytest.raises(capi.YaskError):   # run before prepare (reference: context.cpp:265-266)
        s.run_solution(0, 1)
    with pytest.raises(capi.YaskError):
        capi.Solution("no_such_stencil")
    with pytest.raises(capi.YaskError):
        s.set_option("fp_mode", "7")
    if capi.device_count() == 0:
        with pytest.raises(capi.YaskError) as e:   # no CPU fallback: must fail loudly
            s.prepare_solution(0)
        assert "no CUDA device" in str(e.value)
    s.close()


def test_block_steps_option_and_geometry():
    """The temporal tile's storage is decided before prepare: block_steps >= 2 on iso3dfd radius <= 2 adds a spare pair of slots
    and the pads its boxes reach into; the API-visible step window stays two steps; other radii keep the reference's storage."""
    from yask_b200 import capi
    for radius, bs, want_slots in ((2, 2, 4), (1, 2, 4), (2, 1, 2), (8, 2, 2)):
        s = capi.Solution("iso3dfd", radius=radius)
        s.set_overall_domain_size_vec((64, 48, 96))
        s.set_option("block_steps", bs)
        assert s.get_option("block_steps") == str(bs)
        s.plan_geometry()
        p, v = s.get_var("p").info, s.get_var("v").info
        assert p.step_alloc == 2
        assert p.storage_bytes == want_slots * p.slot_elems * 4
        if want_slots == 4:
            assert p.dims[1].left_pad >= 2 * radius and p.dims[2].left_pad >= 2 * radius and p.dims[3].left_pad >= 8
            assert v.dims[0].left_pad >= radius and v.dims[2].left_pad >= 4
        s.close()


def test_step_to_slot_map_with_spare_slots(tmp_path):
    """tests/unit/slot_test.cpp: Var::slot_of / nslots / valid-step window with the temporal tile's spare pair of slots."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "slot_test")
    subprocess.run(["g++", "-std=c++17", "-I/usr/local/cuda/include", os.path.join(root, "tests", "unit", "slot_test.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr
