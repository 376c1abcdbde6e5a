#!/usr/bin/env python
"""Secondary benchmark lines (BASELINE.json configs 3 and 5): awp_elastic fp32 512^3 and ssg fp64 512^3 per B200
through the C ABI.  Prints one JSON line per stencil with the algorithmic-bytes roofline fraction
(awp_elastic 120 B/point-step, ssg 248 B/point-step).
Under torchrun (WORLD_SIZE > 1) every rank owns n^3 points of an N x 1 x 1 rank grid (weak scaling; halos are pushed
into peer HBM after every stage), the time is the max over ranks and rank 0 prints the whole-job throughput:
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 bench_stencils.py 512"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tests.golden.ranges import RANGES, range_of  # value ranges only: pure data, nothing under oracle/ is imported
from yask_b200 import capi
from yask_b200.synth import var_salt

BYTES = {"awp_elastic": 120, "ssg": 248, "iso3dfd": 16}


RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
_dist = None


def _init_dist():
    global _dist
    if WORLD > 1 and _dist is None:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(LOCAL)
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
        _dist = dist
    return _dist


def run(stencil, n, steps, warm, fp_mode, opts=(), dist=None):
    """One timed run; `dist` = an already initialised torch.distributed (bench.py), else initialised here under torchrun."""
    if dist is None:
        dist = _init_dist()
    s = capi.Solution(stencil, elem_bytes=0)
    s.set_rank_domain_size_vec((n, n, n))
    if WORLD > 1:
        s.set_num_ranks_vec([WORLD, 1, 1])
        s.set_rank_index_vec([RANK, 0, 0])
    s.set_option("fp_mode", fp_mode)
    for kv in opts:
        k, v = kv.split("=", 1)
        s.set_option(k, v)
    s.prepare_solution(LOCAL)
    if WORLD > 1:
        from yask_b200 import multi
        multi.connect(s, dist, RANK, WORLD)
    for v in s.get_vars():
        vi = v.info
        lo, hi = range_of(RANGES[stencil], vi.name.decode())
        for t in (range(vi.step_alloc) if vi.has_step else [0]):
            v.fill_hash(t, 3, var_salt(vi.name.decode(), t), lo, hi)
    s.run_solution(0, warm - 1)
    s.sync()
    if dist:
        dist.barrier()
    s.clear_stats()
    s.run_solution(warm, warm + steps - 1)
    st = s.get_stats()
    secs = st.elapsed_secs
    if dist:
        import torch
        tt = torch.tensor([secs], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)      # device time, max over ranks
        secs = tt.item()
        dist.barrier()
    s.close()
    gpts = n ** 3 * WORLD * steps / secs / 1e9
    peak = 6567.4
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    return {"stencil": stencil, "n": n, "n_gpus": WORLD, "scaling": "weak", "steps": steps, "fp_mode": fp_mode, "gpoints_per_s": round(gpts, 2),
            "ms_per_step": round(secs / steps * 1e3, 4), "algorithmic_gbs": round(gpts * BYTES[stencil], 1),
            "roofline_frac_of_measured_hbm": round(gpts * BYTES[stencil] / peak / WORLD, 4), "kernel_launches": st.kernel_launches,
            "options": list(opts)}


if __name__ == "__main__":
    # usage: bench_stencils.py [n] [key=value engine options ...]
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    opts = [a for a in sys.argv[2:] if "=" in a]
    modes = (2,) if opts else (2, 0)
    for stencil in ("awp_elastic", "ssg"):
        for mode in modes:
            line = run(stencil, n, 10, 3, mode, opts)
            if RANK == 0:
                print(json.dumps(line), flush=True)
    if _dist is not None:
        _dist.destroy_process_group()
