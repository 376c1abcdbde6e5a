"""Scratch tuner: time iso3dfd variants on one GPU (not part of the product; used through gpurun)."""
import json, sys, time
import ctypes as C
sys.path.insert(0, '.'); sys.path.insert(0, '..')
from yask_b200 import capi
from yask_b200.synth import var_salt
import torch

def run(n, steps, warm, **opts):
    s = capi.Solution("iso3dfd")
    s.set_overall_domain_size_vec(n)
    for k, v in opts.items():
        s.set_option(k, v)
    s.prepare_solution(0)
    p, v = s.get_var("p"), s.get_var("v")
    for t in (0, 1):
        p.fill_hash(t, 1, var_salt("p", t), -1.0, 1.0)
    v.fill_hash(0, 1, var_salt("v", 0), 0.05, 0.3)
    s.run_solution(0, warm - 1)
    s.sync(); s.clear_stats()
    s.run_solution(warm, warm + steps - 1)
    st = s.get_stats()
    pts = n[0]*n[1]*n[2]*steps
    gpts = pts / st.elapsed_secs / 1e9
    cs = p.checksum(p.get_last_valid_step_index())
    s.close()
    return gpts, st.elapsed_secs/steps*1e3, cs

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n = (N, N, N)
    res = []
    base = None
    variants = []
    for rep in range(2):
        for (cs, pc, ph) in ((0, 0, 0), (1, 0, 0), (1, 2, 0), (1, 2, 2), (0, 2, 2), (1, 0, 2)):
            variants.append(dict(kernel="tma", tile=7, st_cs=cs, pol_c=pc, pol_h=ph))
    for opts in variants:
        steps = 3 if opts.get("kernel") == "direct" else 40
        g, ms, cs = run(n, steps, 3, **opts)
        print(json.dumps(dict(opts=opts, gpts=round(g, 2), ms_per_step=round(ms, 4), gbps=round(g*16, 1), checksum=cs)), flush=True)
