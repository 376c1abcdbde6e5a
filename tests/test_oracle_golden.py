"""CPU: pin the oracle (oracle/yask_oracle.c) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py via oracle/ref_driver.cpp)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import contract_mode_of, golden_cases, load_golden, regen_inputs


def test_fd_coefficients_match_generated_literals():
    # Literals as printed into the reference's generated kernel for iso3dfd radius 8
    # (oracle/_ref/.../gen/yask_stencil_code.hpp, emitted by src/compiler/lib/Cpp.cpp:39-53).
    lit = ["-3.665812925170066e-03", "7.111111111111103e-04", "-1.244444444444443e-04", "3.016835016835014e-05",
           "-7.070707070707062e-06", "1.392385392385389e-06", "-2.072002072002067e-07", "2.029716315430595e-08",
           "-9.712509712509679e-10"]
    c = O.iso3dfd_coeffs(8)
    assert [float(s) for s in lit] == list(c)
    # radius 2 literals from the "-target pseudo" listing in SURVEY.md Appendix A (6 digits).
    c2 = O.iso3dfd_coeffs(2)
    assert abs(c2[0] - (-3.0e-3)) < 1e-9 and abs(c2[1] - 0.000533333) < 1e-9 and abs(c2[2] - (-3.333333e-05)) < 1e-11


@pytest.mark.parametrize("path", golden_cases("iso3dfd_avx512") + golden_cases("iso3dfd-strict"))
def test_iso3dfd_oracle_bit_exact_vs_reference(path):
    meta, arrays = load_golden(path)
    ins = regen_inputs(meta)
    mode = contract_mode_of(meta["ref_tag"])
    out = O.iso3dfd_run(ins[("p", 0)], ins[("p", 1)], ins[("v", 0)], 8, meta["steps"], mode)
    t_last = meta["vars"]["p"]["steps"][1]
    ref = arrays[f"p.t{t_last}"]
    got = out[8:-8, 8:-8, 8:-8]
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # and the other contraction mode must NOT be what this build does (guards the fixture labels)
    if meta["steps"] >= 2:
        other = O.iso3dfd_run(ins[("p", 0)], ins[("p", 1)], ins[("v", 0)], 8, meta["steps"], 2 - mode)
        assert not np.array_equal(other[8:-8, 8:-8, 8:-8].view(np.uint32), ref.view(np.uint32))


# ---- emitter-generated solutions (awp_elastic fp32, ssg fp64) ---------------------------------------
from tests.helpers import field_ulps  # noqa: E402


def _domain(ir, name, arr):
    v = [x for x in ir["vars"] if x["name"] == name][0]
    vd = [d for d in v["dims"] if d != ir["step_dim"]]
    return arr[tuple(slice(v["halo"][d][0], arr.shape[i] - v["halo"][d][1]) if d in v["halo"] else slice(None) for i, d in enumerate(vd))]


from tests.helpers import generated_golden_cases  # noqa: E402


# Solutions whose default-build contraction is not reproduced element for element: held to 4 field-ulps instead.
NOT_BITEXACT_DEFAULT = {"tti"}


@pytest.mark.parametrize("path", generated_golden_cases())
def test_generated_oracle_vs_reference(path):
    """Strict reference build (-ffp-contract=off): bit-exact.  Default build (GCC -O3, -ffp-contract=fast): bit-exact as
    well, with the products fused exactly where GCC fuses them (emitter: contract_like_gcc; oracle contract=1) -- for 44 of
    the 45 solutions; tti (NOT_BITEXACT_DEFAULT) within 4 field-ulps."""
    meta, arrays = load_golden(path)
    ins = regen_inputs(meta)
    strict = "strict" in meta["ref_tag"]
    out = O.gen_run(meta["stencil"], meta["n"], meta["steps"], ins, contract=0 if strict else 1)
    ir = O.gen_ir(meta["stencil"])
    assert len(out) == sum(1 for v in ir["vars"] if v["is_output"])
    for name, (tl, arr) in out.items():
        ref = arrays[f"{name}.t{tl}"]
        got = _domain(ir, name, arr)
        assert got.shape == ref.shape and got.dtype == ref.dtype
        if strict or meta["stencil"] not in NOT_BITEXACT_DEFAULT:
            it = np.uint32 if got.dtype == np.float32 else np.uint64
            assert np.array_equal(got.view(it), ref.view(it)), name
        else:
            assert field_ulps(got, ref) <= 4.0, (name, field_ulps(got, ref))


@pytest.mark.parametrize("path", golden_cases("iso3dfd-r"))
def test_iso3dfd_small_radius_oracle_bit_exact_vs_reference(path):
    """iso3dfd built by the reference at radius 1 and 2 (`make stencil=iso3dfd radius=<r>`, both FP builds): the radii of the
    temporal tile.  Pins the oracle -- and through it the CTA emulator and the GPU tests of the tile -- to the unmodified
    reference at these radii too."""
    meta, arrays = load_golden(path)
    radius = int(meta["ref_tag"].split("-r")[1][0])
    ins = regen_inputs(meta)
    mode = contract_mode_of(meta["ref_tag"])
    out = O.iso3dfd_run(ins[("p", 0)], ins[("p", 1)], ins[("v", 0)], radius, meta["steps"], mode)
    t_last = meta["vars"]["p"]["steps"][1]
    ref = arrays[f"p.t{t_last}"]
    got = out[radius:-radius, radius:-radius, radius:-radius]
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
