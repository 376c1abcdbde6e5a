#!/usr/bin/env python
"""Temporal-tile benchmark line: iso3dfd radius 2 (4th order in space) fp32, n^3 points on one B200, the one-step sweep kernel
against the two-steps-per-sweep temporal tile (option block_steps=2, the reference's -bt 2) through the C ABI, same inputs.
Prints ONE JSON line: ms per step and GPts/s of both paths, whether their results are bit-identical (checksums of the last two
steps), and the roofline reading -- per point and step the one-step sweep moves 16 B, the temporal tile 10 B (20 B per pair of
steps: read p(t-1), p(t), v; write p(t+1), p(t+2)); `frac_16B` above 1 is therefore possible for the tile and is the point of it.
bench.py runs this file in a subprocess (a new kernel must not be able to take the headline line down with it)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from yask_b200 import capi
from yask_b200.synth import var_salt


def run(radius, n, block_steps, warm, steps, variant=-1):
    s = capi.Solution("iso3dfd", radius=radius)
    s.set_overall_domain_size_vec((n, n, n))
    s.set_option("block_steps", block_steps)
    s.set_option("tt_variant", variant)
    s.prepare_solution(0)
    p, v = s.get_var("p"), s.get_var("v")
    for t in (0, 1):
        p.fill_hash(t, 7, var_salt("p", t), -1.0, 1.0)
    v.fill_hash(0, 7, var_salt("v", 0), 0.05, 0.3)
    s.run_solution(0, warm - 1)
    s.sync()
    s.clear_stats()
    s.run_solution(warm, warm + steps - 1)
    st = s.get_stats()
    tl = p.get_last_valid_step_index()
    cs = [str(p.checksum(tl)), str(p.checksum(tl - 1))]
    s.close()
    ms = st.elapsed_secs * 1e3 / steps
    return {"ms_per_step": round(ms, 4), "gpoints_per_s": round(n ** 3 / ms / 1e6, 2), "kernel_launches": int(st.kernel_launches), "checksums": cs}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    radius = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    peak = 6567.4
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    one = run(radius, n, 1, 4, steps)
    # the form measured on a B200 first; then the others, each in its own try (a CUDA fault ends the process's context: what was
    # measured before it is still reported)
    names = ["shared_memory", "x_queues", "shared_memory_512_threads", "x_queues_512_threads"]
    forms, errors = {}, {}
    for vi, nm in enumerate(names):
        try:
            forms[nm] = run(radius, n, 2, 4, steps, vi)
            forms[nm]["bit_identical"] = forms[nm]["checksums"] == one["checksums"]
        except Exception as e:
            errors[nm] = repr(e)[:200]
    try:
        tt = run(radius, n, 2, 4, steps)          # the engine's own choice of form
    except Exception as e:
        errors["default"] = repr(e)[:200]
        tt = dict(forms["shared_memory"])
    for r, bpp in ((one, 16), (tt, 10)):
        r["algorithmic_bytes_per_point_step"] = bpp
        r["achieved_gbs"] = round(r["gpoints_per_s"] * bpp, 1)
        r["frac"] = round(r["gpoints_per_s"] * bpp / peak, 4)
    tt["frac_16B"] = round(tt["gpoints_per_s"] * 16 / peak, 4)
    print(json.dumps({"workload": f"iso3dfd radius {radius} fp32, {n}^3 points, 1 GPU: one-step sweep vs temporal tile (2 steps per sweep)",
                      "steps": steps, "warmup": 4, "peak_gbs": peak, "one_step": one, "temporal_tile": tt,
                      "forms": {k: {"ms_per_step": v["ms_per_step"], "gpoints_per_s": v["gpoints_per_s"], "bit_identical": v["bit_identical"]} for k, v in forms.items()},
                      "form_errors": errors,
                      "bit_identical": one["checksums"] == tt["checksums"], "speedup": round(one["ms_per_step"] / tt["ms_per_step"], 4)}), flush=True)
