"""The CUDA emitter (yask_b200/emitter/yask_cuda_emit.py) re-run on the reference compiler's output must reproduce the
committed generated sources (kernels + oracle restatement) -- build container only (needs tools/_refc/bin/yask_compiler.exe, built by oracle/build_ref.sh);
plus parser unit checks that need no reference."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "yask_b200", "emitter"))
import yask_cuda_emit as E  # noqa: E402

GEN = os.path.join(ROOT, "yask_b200", "csrc", "gen")
MAN = json.load(open(os.path.join(GEN, "manifest.json")))


@pytest.mark.parametrize("name", ["awp_elastic", "ssg", "test_scratch_2d", "test_scratch_boundary_1d", "test_step_cond_1d", "test_func_1d",
                                  "test_reverse_2d", "iso3dfd_sponge", "wave2d"])
def test_emitter_reproduces_committed_sources(name):
    if not os.path.exists(E.COMPILER):
        pytest.skip("reference compiler not built (oracle/build_ref.sh; build container only)")
    m = MAN[name]
    text = E.run_frontend(m["stencil"], m["elem_bytes"], m.get("radius") or None)
    ir = E.parse_generated(text, name)
    ident = E.c_ident(name)
    assert E.emit_cuda(ir) == open(os.path.join(GEN, f"{ident}.gen.cuh")).read()
    assert E.emit_oracle(ir) == open(os.path.join(ROOT, "oracle", "gen", f"{ident}.gen.h")).read()


def test_manifest_matches_registry_and_json():
    inc = open(os.path.join(GEN, "gen_all.inc")).read()
    for name in MAN:
        assert f'{{"{name}", yb::gen::{E.c_ident(name)}_register}}' in inc
        ir = json.load(open(os.path.join(GEN, f"{name}.json")))
        idx = sorted(p["index"] for st in ir["stages"] for p in st["parts"])
        assert len(set(idx)) >= 1 and all(p["name"] for st in ir["stages"] for p in st["parts"])
        # scratch parts precede the part that needs them; scratch vars come last in the var table
        flags = [bool(v.get("scratch")) for v in ir["vars"]]
        assert flags == sorted(flags)
        for st in ir["stages"]:
            assert not st["parts"][-1]["scratch"]


def test_expression_parser_keeps_evaluation_order():
    t = E.Parser("expr_temp3 + @0 * 2.5 - (expr_temp4 / @1)").parse()
    assert t == ("sub", ("add", ("tmp", "e3"), ("mul", ("read", 0), ("const", "2.5"))), ("div", ("tmp", "e4"), ("read", 1)))
    t = E.Parser("yask_max(arg0_temp3, -1.5e+00) * yask_cbrt(@0)").parse()
    assert t == ("mul", ("call", "max", [("tmp", "a0_3"), ("const", "-1.5e+00")]), ("call", "cbrt", [("read", 0)]))
    with pytest.raises(E.EmitError):
        E.Parser("yask_tanh(@0)").parse()


def test_domain_condition_bounds():
    c = E.parse_domain_cond("((x >= (FIRST_INDEX(x) + 5)) && (x <= (LAST_INDEX(x) - 3)))", ["x"])
    assert c["bounds"] == {"2": [["GF", 5], ["GL", -3]]} and "G(2)" in c["expr"]
    c = E.parse_domain_cond("((x < (FIRST_INDEX(x) + 20)) || (y > 3))", ["x", "y"])
    assert c["bounds"] is None
