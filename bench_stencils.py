#!/usr/bin/env python
"""Secondary benchmark lines (BASELINE.json configs 3 and 5): awp_elastic fp32 512^3 and ssg fp64 512^3 on one
B200 through the C ABI.  Prints one JSON line per stencil with the algorithmic-bytes roofline fraction
(awp_elastic 120 B/point-step, ssg 248 B/point-step)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tests.golden.make_golden import RANGES, range_of  # value ranges only  (no oracle code is executed)
from yask_b200 import capi
from yask_b200.synth import var_salt

BYTES = {"awp_elastic": 120, "ssg": 248, "iso3dfd": 16}


def run(stencil, n, steps, warm, fp_mode, opts=()):
    s = capi.Solution(stencil, elem_bytes=0)
    s.set_overall_domain_size_vec((n, n, n))
    s.set_option("fp_mode", fp_mode)
    for kv in opts:
        k, v = kv.split("=", 1)
        s.set_option(k, v)
    s.prepare_solution(0)
    for v in s.get_vars():
        vi = v.info
        lo, hi = range_of(RANGES[stencil], vi.name.decode())
        for t in (range(vi.step_alloc) if vi.has_step else [0]):
            v.fill_hash(t, 3, var_salt(vi.name.decode(), t), lo, hi)
    s.run_solution(0, warm - 1)
    s.sync()
    s.clear_stats()
    s.run_solution(warm, warm + steps - 1)
    st = s.get_stats()
    s.close()
    gpts = n ** 3 * steps / st.elapsed_secs / 1e9
    peak = 6567.4
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    return {"stencil": stencil, "n": n, "steps": steps, "fp_mode": fp_mode, "gpoints_per_s": round(gpts, 2),
            "ms_per_step": round(st.elapsed_secs / steps * 1e3, 4), "algorithmic_gbs": round(gpts * BYTES[stencil], 1),
            "roofline_frac_of_measured_hbm": round(gpts * BYTES[stencil] / peak, 4), "kernel_launches": st.kernel_launches, "options": list(opts)}


if __name__ == "__main__":
    # usage: bench_stencils.py [n] [key=value engine options ...]
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    opts = [a for a in sys.argv[2:] if "=" in a]
    modes = (2,) if opts else (2, 0)
    for stencil in ("awp_elastic", "ssg"):
        for mode in modes:
            print(json.dumps(run(stencil, n, 10, 3, mode, opts)), flush=True)
