"""iso3dfd temporal tile on the GPU (option block_steps / -bt 2, radius <= 2; yask_b200/csrc/yb_iso3dfd_tt.cuh) through the
C ABI: temporal blocking must never change a result (/root/reference/src/kernel/lib/context.cpp:657-681) -- bit-exact against
the oracle, against the one-step kernels, for odd step counts (the left-over step takes the one-step path), runs split over
several run_solution() calls, ragged domains, every FP mode; halo cells and the API's two-step window behave as without it."""
import numpy as np
import pytest

from oracle import oracle as O
from yask_b200 import capi
from yask_b200.synth import hash_field, var_salt

pytestmark = pytest.mark.gpu


def synth(n, seed, R):
    return {("p", 0): hash_field(seed, var_salt("p", 0), (-R, -R, -R), [i + 2 * R for i in n], -1, 1),
            ("p", 1): hash_field(seed, var_salt("p", 1), (-R, -R, -R), [i + 2 * R for i in n], -1, 1),
            ("v", 0): hash_field(seed, var_salt("v", 0), (0, 0, 0), n, 0.05, 0.3)}


def make(n, R, ins, block_steps, fp_mode=2, opts=None):
    s = capi.Solution("iso3dfd", radius=R)
    s.set_overall_domain_size_vec(n)
    s.set_option("fp_mode", fp_mode)
    s.set_option("block_steps", block_steps)
    for k, v in (opts or {}).items():
        s.set_option(k, v)
    s.prepare_solution(0)
    p, v = s.get_var("p"), s.get_var("v")
    for t in (0, 1):
        p.set_elements_in_slice(ins[("p", t)], *p.halo_box(t))
    v.set_elements_in_slice(ins[("v", 0)], *v.halo_box(0))
    return s


def result(s):
    p = s.get_var("p")
    tl = p.get_last_valid_step_index()
    return p.get_elements_in_slice(*p.domain_box(tl)), p.get_elements_in_slice(*p.domain_box(tl - 1))


@pytest.mark.parametrize("variant", [0, 1, 2, 3])     # 0: neighbours from shared memory, 1: x neighbours in register queues; 2, 3: the same with 512 threads
@pytest.mark.parametrize("fp_mode", [2, 0])
@pytest.mark.parametrize("steps", [2, 3, 4, 7])
@pytest.mark.parametrize("R,n", [(2, (40, 37, 150)), (1, (33, 20, 260)), (2, (9, 16, 128)), (1, (64, 48, 64))])
def test_temporal_tile_bit_exact_vs_oracle(R, n, steps, fp_mode, variant):
    ins = synth(n, 21, R)
    s = make(n, R, ins, 2, fp_mode, {"tt_variant": variant})
    s.run_solution(0, steps - 1)
    got, got_prev = result(s)
    st = s.get_stats()
    s.close()
    ref = O.iso3dfd_run(ins[("p", 0)], ins[("p", 1)], ins[("v", 0)], R, steps, fp_mode)[R:-R, R:-R, R:-R]
    ref_prev = O.iso3dfd_run(ins[("p", 0)], ins[("p", 1)], ins[("v", 0)], R, steps - 1, fp_mode)[R:-R, R:-R, R:-R]
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))
    assert np.array_equal(got_prev.view(np.uint32), np.ascontiguousarray(ref_prev).view(np.uint32))
    # the fused path really ran: one launch per pair of steps + one for a left-over step
    assert st.kernel_launches == steps // 2 + steps % 2
    assert st.num_steps_done == steps


def test_split_runs_halo_cells_and_step_window():
    """run(0,1); run(2,4) == run(0,4); halo cells keep what the user wrote; the API window stays two steps long."""
    R, n = 2, (24, 21, 140)
    ins = synth(n, 5, R)
    a = make(n, R, ins, 2)
    a.run_solution(0, 1)
    a.run_solution(2, 4)
    b = make(n, R, ins, 1)
    b.run_solution(0, 4)
    for x, y in zip(result(a), result(b)):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    pa, pb = a.get_var("p"), b.get_var("p")
    assert pa.info.step_alloc == 2 and pa.get_first_valid_step_index() == pb.get_first_valid_step_index()
    tl = pa.get_last_valid_step_index()
    for t in (tl - 1, tl):      # whole halo boxes, halo cells included, equal those of the two-slot run
        assert np.array_equal(pa.get_elements_in_slice(*pa.halo_box(t)).view(np.uint32), pb.get_elements_in_slice(*pb.halo_box(t)).view(np.uint32))
    a.close()
    b.close()


@pytest.mark.parametrize("R", [1, 2])
def test_temporal_equals_one_step_kernels_large(R):
    """Size-independent property at a size the oracle cannot reach: same bits as the one-step sweep kernel (checksum)."""
    n, steps = (200, 150, 300), 6
    sums = []
    for bs, variant in ((2, 0), (2, 1), (2, 2), (2, 3), (1, 0)):
        s = capi.Solution("iso3dfd", radius=R)
        s.set_overall_domain_size_vec(n)
        s.set_option("block_steps", bs)
        s.set_option("tt_variant", variant)
        s.prepare_solution(0)
        p, v = s.get_var("p"), s.get_var("v")
        for t in (0, 1):
            p.fill_hash(t, 7, var_salt("p", t), -1.0, 1.0)
        v.fill_hash(0, 7, var_salt("v", 0), 0.05, 0.3)
        s.run_solution(0, steps - 1)
        tl = p.get_last_valid_step_index()
        sums.append((p.checksum(tl), p.checksum(tl - 1)))
        s.close()
    assert all(x == sums[-1] for x in sums[:-1])


def test_offline_tuner_chooses_between_one_and_two_steps_per_sweep():
    """run_auto_tuner_now on a solution with a temporal tile times one step per sweep against both compiled two-step forms and keeps
    the fastest (aux/yk_solution_api.hpp:858-882); whatever it keeps, results stay exact."""
    R, n = 2, (96, 64, 256)
    ins = synth(n, 9, R)
    s = make(n, R, ins, 2)
    rep = s.run_auto_tuner_now()
    assert "one step per sweep" in rep and "x queues" in rep and "best:" in rep
    assert s.get_option("block_steps") in ("1", "2")
    p, v = s.get_var("p"), s.get_var("v")
    for t in (0, 1):       # the tuner does not preserve var contents
        p.set_elements_in_slice(ins[("p", t)], *p.halo_box(t))
    s.run_solution(0, 3)
    got, _ = result(s)
    s.close()
    ref = O.iso3dfd_run(ins[("p", 0)], ins[("p", 1)], ins[("v", 0)], R, 4, 2)[R:-R, R:-R, R:-R]
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))


def test_block_steps_is_inert_where_no_tile_exists():
    """-bt on a radius without a temporal tile (8) or after prepare: accepted, results unchanged."""
    n, R = (32, 24, 64), 8
    ins = synth(n, 3, R)
    s = make(n, R, ins, 2)
    assert s.get_var("p").info.storage_bytes == 2 * s.get_var("p").info.slot_elems * 4
    s.run_solution(0, 1)
    got, _ = result(s)
    s.close()
    ref = O.iso3dfd_run(ins[("p", 0)], ins[("p", 1)], ins[("v", 0)], R, 2, 2)[R:-R, R:-R, R:-R]
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))


from tests.helpers import contract_mode_of, golden_cases, load_golden, regen_inputs  # noqa: E402


def _radius_of(path):
    return int(path.split("iso3dfd-r")[1][0])


@pytest.mark.parametrize("kernel", ["tma", "direct"])
@pytest.mark.parametrize("path", golden_cases("iso3dfd-r"))
def test_one_step_kernels_bit_exact_vs_reference_fixture_at_small_radii(path, kernel):
    """The one-step sweep kernel of each radius and the direct kernel against the unmodified reference built at radius 1, 2, 4."""
    meta, arrays = load_golden(path)
    R = _radius_of(path)
    ins = regen_inputs(meta)
    s = make(meta["n"], R, ins, 1, contract_mode_of(meta["ref_tag"]), {"kernel": kernel})
    s.run_solution(0, meta["steps"] - 1)
    got, _ = result(s)
    s.close()
    ref = arrays[f"p.t{meta['vars']['p']['steps'][1]}"]
    assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("path", [p for p in golden_cases("iso3dfd-r") if _radius_of(p) <= 2])
def test_temporal_tile_bit_exact_vs_reference_fixture(path, variant):
    """Against the outputs of the UNMODIFIED reference built at radius 1 / 2 (`make stencil=iso3dfd radius=<r>`), default GCC build
    (fp_mode 2) and -ffp-contract=off build (fp_mode 0): 4 steps = two fused launches, 5 steps = two fused + one one-step launch."""
    meta, arrays = load_golden(path)
    R = int(meta["ref_tag"].split("-r")[1][0])
    ins = regen_inputs(meta)
    s = make(meta["n"], R, ins, 2, contract_mode_of(meta["ref_tag"]), {"tt_variant": variant})
    s.run_solution(0, meta["steps"] - 1)
    got, _ = result(s)
    st = s.get_stats()
    s.close()
    ref = arrays[f"p.t{meta['vars']['p']['steps'][1]}"]
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert st.kernel_launches == meta["steps"] // 2 + meta["steps"] % 2
