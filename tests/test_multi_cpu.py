"""CPU (gloo, world_size 2): host-side logic of the multi-rank path -- rank geometry planning through
the C ABI, neighbour enumeration, and the blob rendezvous plumbing of yask_b200.multi."""
import os

import pytest
import torch.multiprocessing as mp

from yask_b200 import capi, multi


def test_grid_coords_roundtrip_and_neighbours():
    g = (2, 3, 4)
    for r in range(24):
        assert multi.linear_rank(multi.grid_coords(r, g), g) == r
    assert multi.grid_coords(5, (8, 1, 1)) == [5, 0, 0]
    assert sorted(multi.neighbours(0, (2, 2, 2))) == [1, 2, 3, 4, 5, 6, 7]
    assert sorted(multi.neighbours(3, (8, 1, 1))) == [2, 4]
    assert multi.neighbours(0, (1, 1, 1)) == []


@pytest.mark.parametrize("overall,grid", [((1000, 64, 64), (3, 1, 1)), ((128, 130, 70), (2, 2, 2)), ((8192, 1024, 1024), (8, 1, 1))])
def test_rank_geometry_tiles_the_domain(overall, grid):
    """Reference rule (setup.cpp:462-503): ceil(overall/nranks) per rank, remainder on the last."""
    world = grid[0] * grid[1] * grid[2]
    cover = [[0] * o for o in (overall[0], overall[1], overall[2])] if max(overall) <= 2048 else None
    total = 0
    for r in range(world):
        s = capi.Solution("iso3dfd")
        s.set_overall_domain_size_vec(overall)
        s.set_num_ranks_vec(grid)
        idx = multi.grid_coords(r, grid)
        s.set_rank_index_vec(idx)
        s.plan_geometry()
        size = s.get_rank_domain_size_vec()
        first = s.get_first_rank_domain_index_vec()
        for d in range(3):
            per = -(-overall[d] // grid[d])
            assert first[d] == per * idx[d]
            assert size[d] == (per if idx[d] < grid[d] - 1 else overall[d] - per * (grid[d] - 1))
        total += size[0] * size[1] * size[2]
        # var geometry is planned too: p carries the halo, v does not
        vi = s.get_var("p").info
        assert vi.dims[1].rank_offset == first[0] and vi.dims[1].left_pad >= 8 and vi.dims[3].left_pad % 32 == 0
        s.close()
    assert total == overall[0] * overall[1] * overall[2]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = capi.Solution("iso3dfd")
    s.set_overall_domain_size_vec((100, 32, 32))
    s.set_num_ranks_vec((world, 1, 1))
    s.set_rank_index_vec(multi.grid_coords(rank, (world, 1, 1)))
    s.plan_geometry()
    mine = (rank, s.get_first_rank_domain_index_vec(), s.get_rank_domain_size_vec())
    got = [None] * world
    dist.all_gather_object(got, mine)      # the same collective multi.connect() uses for the IPC blobs
    err = None
    try:                                   # no device here: exporting a blob must fail loudly, not fake it
        s.halo_export()
    except capi.YaskError as e:
        err = str(e)
    q.put((rank, got, err))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_rendezvous_and_geometry():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, err in res:
        assert [g[0] for g in got] == [0, 1]
        assert got[0][1] == [0, 0, 0] and got[0][2] == [50, 32, 32]
        assert got[1][1] == [50, 0, 0] and got[1][2] == [50, 32, 32]
        assert err is not None and "not prepared" in err


def _comm_worker(rank, world, key, q):
    import ctypes as C
    import numpy as np
    L = capi.lib()
    capi._chk(L.yb_comm_init(rank, world, key.encode()))
    assert L.yb_comm_rank() == rank and L.yb_comm_world() == world
    # all-gather larger than one mailbox slot (64 KiB): exercises the chunked path
    n = 100_000
    mine = (np.arange(n, dtype=np.int32) * (rank + 1)).astype(np.int32)
    allv = np.zeros(world * n, np.int32)
    capi._chk(L.yb_comm_allgather(mine.ctypes.data_as(C.c_void_p), mine.nbytes, allv.ctypes.data_as(C.c_void_p)))
    ok = all(np.array_equal(allv[r * n:(r + 1) * n], np.arange(n, dtype=np.int32) * (r + 1)) for r in range(world))
    s = C.c_int64(0)
    capi._chk(L.yb_comm_sum_i64(10 + rank, C.byref(s)))
    m = C.c_double(0)
    capi._chk(L.yb_comm_max_f64(1.5 * rank, C.byref(m)))
    for _ in range(50):
        capi._chk(L.yb_comm_barrier())
    L.yb_comm_finalize()
    q.put((rank, ok, s.value, m.value))


def test_shared_memory_rendezvous_three_processes():
    """The library's own communicator (yb_comm.cpp: /dev/shm mailbox, no MPI, no torch): all-gather, integer sum, max and
    barriers between three processes -- what yk_env::global_barrier / sum_over_ranks and the halo-blob exchange of
    prepare_solution() run on."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    key = f"pytest_{os.getpid()}"
    procs = [ctx.Process(target=_comm_worker, args=(r, world, key, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for r, (rank, ok, s, m) in enumerate(res):
        assert rank == r and ok and s == sum(10 + i for i in range(world)) and m == 1.5 * (world - 1)
