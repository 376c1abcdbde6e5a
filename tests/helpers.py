"""Shared test helpers: golden-fixture loading and synthetic-input regeneration."""
import glob
import json
import os

import numpy as np

from yask_b200.synth import hash_field, var_salt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def generated_golden_cases():
    """Fixtures of every emitter-generated solution (yask_b200/csrc/gen/manifest.json)."""
    man = json.load(open(os.path.join(os.path.dirname(GOLDEN), "..", "yask_b200", "csrc", "gen", "manifest.json")))
    out = []
    for p in golden_cases(""):
        st = json.loads(str(np.load(p)["meta"]))["stencil"]
        if st in man:
            out.append(p)
    return out


def load_golden(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    arrays = {k: z[k] for k in z.files if k != "meta"}
    return meta, arrays


def range_of(ranges, name):
    """Value range of a var: longest matching name prefix (same rule as tests/golden/make_golden.py)."""
    best = None
    for k, v in ranges.items():
        if name.startswith(k) and (best is None or len(k) > len(best[0])):
            best = (k, v)
    return best[1]


def regen_inputs(meta, dtype=None):
    """Recreate the hash-field inputs a fixture was generated from: {(var, api_step): array over in-box}."""
    if dtype is None:
        dtype = np.float64 if "fp64" in meta["ref_tag"] else np.float32
    ins = {}
    for name, g in meta["vars"].items():
        lo, hi = range_of(meta["ranges"], name)
        has_step = len(g["dims"]) == len(g["in_first"]) + 1      # the boxes cover the non-step dims
        t0, t1 = meta["vars_before"][name]["steps"] if has_step else (0, 0)
        shape = [l - f + 1 for f, l in zip(g["in_first"], g["in_last"])]
        for t in range(t0, t1 + 1):
            if shape:
                ins[(name, t)] = hash_field(meta["seed"], var_salt(name, t), g["in_first"], shape, lo, hi, dtype)
            else:
                ins[(name, t)] = np.array(hash_field(meta["seed"], var_salt(name, t), (0,), (1,), lo, hi, dtype)[0], dtype=dtype)
    return ins


def field_ulps(got, ref):
    """max |got - ref| in units of eps * max|ref| ("field ulps"): a well-conditioned error measure for
    fields whose individual values pass through zero."""
    eps = np.finfo(ref.dtype).eps
    return float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() / (eps * np.abs(ref).max()))


def contract_mode_of(tag: str) -> int:
    """FP-contraction mode that reproduces a reference build (see oracle/yask_oracle.c header)."""
    return 0 if "strict" in tag else 2


def ulp_diff_f32(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def elem_ulps(got, ref):
    """Element-wise distance in units in the last place (ordered-integer distance of the bit patterns), float32 or float64."""
    if got.dtype == np.float32:
        return ulp_diff_f32(got, ref).astype(np.float64)
    ai, bi = got.view(np.int64).astype(object), ref.view(np.int64).astype(object)     # exact arithmetic on 64-bit patterns
    sign = 0x7FFFFFFFFFFFFFFF
    ai = np.where(ai < 0, -(ai & sign), ai)
    bi = np.where(bi < 0, -(bi & sign), bi)
    return np.abs(ai - bi).astype(np.float64)
