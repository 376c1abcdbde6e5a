#!/usr/bin/env python
"""Element-wise ulp distance between the generated kernels (fp_mode 2 = nvcc contraction) and the reference's DEFAULT GCC
build (golden fixtures), per output var, for the benchmarked multi-var stencils.  north_star states the tolerance as
1 ulp fp32 / 4 ulp fp64 per element; the strict builds are bit-exact, the default builds differ by each compiler's FMA
choices, which this tool quantifies (GPU box: python tools/ulp_hist.py > gpurun_out/ulp_hist.json)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import generated_golden_cases, load_golden, regen_inputs  # noqa: E402
from tests.test_generated_gpu import run_gpu  # noqa: E402


def ulps(a, b):
    if a.dtype == np.float32:
        ai, bi = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
        sign = 0x7FFFFFFF
    else:
        ai, bi = a.view(np.int64).astype(object), b.view(np.int64).astype(object)      # exact arithmetic on 64-bit patterns
        sign = 0x7FFFFFFFFFFFFFFF
    ai = np.where(ai < 0, -(ai & sign), ai)
    bi = np.where(bi < 0, -(bi & sign), bi)
    return np.abs(ai - bi).astype(np.float64)


out = []
for path in generated_golden_cases():
    meta, arrays = load_golden(path)
    if meta["stencil"] not in ("awp_elastic", "ssg", "awp", "iso3dfd_fp64") or "strict" in meta["ref_tag"]:
        continue
    ins = regen_inputs(meta)
    res, _ = run_gpu(meta["stencil"], meta["n"], meta["steps"], ins, 2)
    for name, (tl, got) in sorted(res.items()):
        ref = arrays[f"{name}.t{tl}"]
        u = ulps(got, ref).ravel()
        mag = np.abs(ref).ravel().astype(np.float64)
        big = mag >= mag.max() * 2.0 ** -6        # elements not produced by heavy cancellation
        rec = {"case": os.path.basename(path), "var": name, "n": int(u.size), "max_ulp": float(u.max()), "frac_0": float((u == 0).mean()),
               "frac_le1": float((u <= 1).mean()), "frac_le4": float((u <= 4).mean()), "p999": float(np.quantile(u, 0.999)),
               "max_ulp_noncancelled": float(u[big].max()) if big.any() else 0.0, "frac_noncancelled": float(big.mean())}
        out.append(rec)
        print(json.dumps(rec), flush=True)
