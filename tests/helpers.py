"""Shared test helpers: golden-fixture loading and synthetic-input regeneration."""
import glob
import json
import os

import numpy as np

from yask_b200.synth import hash_field, var_salt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    arrays = {k: z[k] for k in z.files if k != "meta"}
    return meta, arrays


def regen_inputs(meta, dtype=np.float32):
    """Recreate the hash-field inputs a fixture was generated from: {(var, api_step): array over in-box}."""
    ins = {}
    for name, g in meta["vars"].items():
        lo, hi = meta["ranges"][name]
        has_step = g["dims"][0] == "t"
        t0, t1 = meta["vars_before"][name]["steps"] if has_step else (0, 0)
        shape = [l - f + 1 for f, l in zip(g["in_first"], g["in_last"])]
        for t in range(t0, t1 + 1):
            ins[(name, t)] = hash_field(meta["seed"], var_salt(name, t), g["in_first"], shape, lo, hi, dtype)
    return ins


def contract_mode_of(tag: str) -> int:
    """FP-contraction mode that reproduces a reference build (see oracle/yask_oracle.c header)."""
    return 0 if "strict" in tag else 2


def ulp_diff_f32(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)
