"""GPU check of the iso3dfd temporal tile without pytest / torch (fast start on a fresh box): bit-exactness against the oracle
at small sizes, equality with the one-step kernels at 1024^3 (checksums), and GPts/s of both paths at 1024^3.
Writes gpurun_out/tt_check.json after every phase (a call that is cut off still leaves what it measured)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yask_b200 import capi                      # noqa: E402
from yask_b200.synth import hash_field, var_salt  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "tt_check.json")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
res = {"phases": []}


def save():
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)


def small(R, n, steps, fp_mode):
    from oracle import oracle as O
    ins = {("p", t): hash_field(21, var_salt("p", t), (-R, -R, -R), [i + 2 * R for i in n], -1, 1) for t in (0, 1)}
    vv = hash_field(21, var_salt("v", 0), (0, 0, 0), n, 0.05, 0.3)
    s = capi.Solution("iso3dfd", radius=R)
    s.set_overall_domain_size_vec(n)
    s.set_option("fp_mode", fp_mode)
    s.set_option("block_steps", 2)
    s.prepare_solution(0)
    p, v = s.get_var("p"), s.get_var("v")
    for t in (0, 1):
        p.set_elements_in_slice(ins[("p", t)], *p.halo_box(t))
    v.set_elements_in_slice(vv, *v.halo_box(0))
    s.run_solution(0, steps - 1)
    tl = p.get_last_valid_step_index()
    got = p.get_elements_in_slice(*p.domain_box(tl))
    st = s.get_stats()
    s.close()
    ref = np.ascontiguousarray(O.iso3dfd_run(ins[("p", 0)], ins[("p", 1)], vv, R, steps, fp_mode)[R:-R, R:-R, R:-R])
    bad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
    return {"what": "vs oracle", "R": R, "n": n, "steps": steps, "fp_mode": fp_mode, "mismatches": bad, "points": int(got.size),
            "kernel_launches": int(st.kernel_launches), "ok": bad == 0 and st.kernel_launches == steps // 2 + steps % 2}


def big(R, N, bs, warm, steps, opts=None):
    s = capi.Solution("iso3dfd", radius=R)
    s.set_overall_domain_size_vec((N, N, N))
    s.set_option("block_steps", bs)
    for k, v in (opts or {}).items():
        s.set_option(k, v)
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
    cs = [p.checksum(tl), p.checksum(tl - 1)]
    s.close()
    ms = st.elapsed_secs * 1e3 / steps
    return {"what": "1024 perf" if N == 1024 else f"{N} perf", "R": R, "N": N, "block_steps": bs, "opts": opts or {}, "steps": steps, "ms_per_step": ms,
            "gpts": N ** 3 / ms / 1e6, "alg_GBs_16B": 16 * N ** 3 / ms / 1e6, "kernel_launches": int(st.kernel_launches), "checksums": [str(c) for c in cs]}


def phase(fn, *a, **k):
    t0 = time.time()
    try:
        r = fn(*a, **k)
    except Exception as e:      # a CUDA fault poisons the context: record and go on (later phases will say so too)
        r = {"what": fn.__name__, "args": repr(a), "error": str(e)[:300], "ok": False}
    r["wall_s"] = round(time.time() - t0, 2)
    res["phases"].append(r)
    save()
    print(json.dumps(r), flush=True)
    return r


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    phase(small, 2, (40, 37, 150), 5, 2)
    phase(small, 1, (33, 20, 260), 4, 0)
    phase(small, 2, (9, 16, 128), 2, 2)
    a = phase(big, 2, N, 2, 4, 20)
    b = phase(big, 2, N, 1, 4, 20)
    if "checksums" in a and "checksums" in b:
        res["r2_equal_to_one_step_kernels"] = a["checksums"] == b["checksums"]
        res["r2_speedup"] = b["ms_per_step"] / a["ms_per_step"]
    a1 = phase(big, 1, N, 2, 4, 20)
    b1 = phase(big, 1, N, 1, 4, 20)
    if "checksums" in a1 and "checksums" in b1:
        res["r1_equal_to_one_step_kernels"] = a1["checksums"] == b1["checksums"]
        res["r1_speedup"] = b1["ms_per_step"] / a1["ms_per_step"]
    # second sample of the r=2 pair (boxes drift as they warm up) and a chunk-length variant
    phase(big, 2, N, 2, 4, 40)
    phase(big, 2, N, 1, 4, 40)
    phase(big, 2, N, 2, 4, 20, {"lx": 256})
    save()
    print(json.dumps({k: v for k, v in res.items() if k != "phases"}))
