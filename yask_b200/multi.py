"""Multi-GPU launcher glue: one process per GPU (torchrun), torch.distributed only for the rendezvous.

The data path never touches torch or NCCL: after `connect()` every rank holds CUDA-IPC mappings of
its neighbours' var storage and the engine's kernels write halos straight into peer HBM over
NVLink (yask_b200/csrc/yb_halo.cu).  This replaces the reference's MPI rank set-up
(/root/reference/src/kernel/lib/setup.cpp:169-524) and buffer allocation (alloc.cpp:456-1031).
"""
from __future__ import annotations

from typing import Sequence


def grid_coords(rank: int, num_ranks: Sequence[int]) -> list[int]:
    """Rank index vector of a linear rank: row-major over the domain dims (x slowest) -- the same
    convention yb_halo uses for peer ranks."""
    idx = []
    for n in reversed(list(num_ranks)):
        idx.append(rank % n)
        rank //= n
    return list(reversed(idx))


def linear_rank(idx: Sequence[int], num_ranks: Sequence[int]) -> int:
    r = 0
    for i, n in zip(idx, num_ranks):
        r = r * n + i
    return r


def neighbours(rank: int, num_ranks: Sequence[int]) -> list[int]:
    """Linear ranks in the 3^N - 1 neighbourhood (no periodic wrap), as yb_halo_prepare enumerates them."""
    idx = grid_coords(rank, num_ranks)
    out = []
    nd = len(num_ranks)

    def rec(d, cur):
        if d == nd:
            if any(c != i for c, i in zip(cur, idx)):
                out.append(linear_rank(cur, num_ranks))
            return
        for o in (-1, 0, 1):
            c = idx[d] + o
            if 0 <= c < num_ranks[d]:
                rec(d + 1, cur + [c])

    rec(0, [])
    return out


def connect(soln, dist, rank: int, world: int) -> None:
    """Exchange halo blobs (CUDA IPC handles) between all ranks and wire up the neighbours."""
    blob = soln.halo_export()
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    for r, b in enumerate(blobs):
        if r != rank:
            soln.halo_import(r, b)
    soln.halo_finalize()


def connect_local(solns) -> None:
    """Same wiring for several ranks living in ONE process (tests): blobs are passed directly."""
    blobs = [s.halo_export() for s in solns]
    for r, s in enumerate(solns):
        for q, b in enumerate(blobs):
            if q != r:
                s.halo_import(q, b)
        s.halo_finalize()


def connect_shm(soln, rank: int, world: int, job_key: str | None = None) -> None:
    """Same wiring WITHOUT torch or MPI: the library's own shared-memory rendezvous (yask_b200/csrc/yb_comm.cpp), the path
    the C++ yk_solution::prepare_solution() takes when the job has more than one rank."""
    from . import capi
    L = capi.lib()
    capi._chk(L.yb_comm_init(rank, world, job_key.encode() if job_key else None))
    capi._chk(L.yb_halo_connect(soln._h))
