"""Deterministic synthetic inputs keyed on GLOBAL logical indices.

The reference's own `init_diff` initialiser depends on its folded storage layout
(/root/reference/src/kernel/lib/generic_var.cpp:168-186), so it cannot be
reproduced by an engine with a different layout.  Tests, fixtures and the bench therefore
use this generator instead: the value at global index (i0, i1, i2) of a var is a hash of
(seed, var_salt, i0, i1, i2) mapped to [lo, hi).  Because it is a function of the *global*
index, any domain decomposition sees the same global field.

The same function is implemented on the device in yask_b200/csrc/yb_fill.cu
(`yb_var_fill_hash`); tests check the two agree bit-for-bit.
"""
from __future__ import annotations

import numpy as np

_OFF = np.uint64(1 << 19)  # index bias so that halo (negative) indices are positive keys
_M = np.uint64(0xFFFFF)    # 20 bits per index


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def var_salt(name: str, step: int = 0) -> int:
    """Stable 32-bit salt from a var name and API step index (FNV-1a)."""
    h = 0x811C9DC5
    for ch in (name + "#" + str(int(step))).encode():
        h = ((h ^ ch) * 0x01000193) & 0xFFFFFFFF
    return h


def hash_field(seed: int, salt: int, first, shape, lo: float, hi: float, dtype=np.float32) -> np.ndarray:
    """Field over the box starting at global index `first` (len<=3) with `shape`.

    u = top 24 bits of splitmix64(key) * 2^-24, value = dtype(lo + (hi - lo) * u) with the
    affine map evaluated in float64.
    """
    first = list(first)
    shape = list(shape)
    nd = len(shape)
    if nd == 4:   # 4-D var (e.g. x,y,z + a misc dim): the leading index perturbs the salt, the rest is a 3-D field
        return np.stack([hash_field(seed, (salt + (first[0] + m) * 0x9E3779B1) & 0xFFFFFFFF, first[1:], shape[1:], lo, hi, dtype)
                         for m in range(shape[0])])
    assert nd <= 3 and len(first) == nd
    while len(shape) < 3:  # left-pad missing dims with index 0
        shape.insert(0, 1)
        first.insert(0, 0)
    idx = []
    for d in range(3):
        a = (np.arange(shape[d], dtype=np.int64) + np.int64(first[d])).astype(np.uint64)
        with np.errstate(over="ignore"):
            a = (a + _OFF) & _M
        idx.append(a)
    key = (idx[0][:, None, None] << np.uint64(40)) | (idx[1][None, :, None] << np.uint64(20)) | idx[2][None, None, :]
    with np.errstate(over="ignore"):
        key = key ^ (np.uint64(seed & 0xFFFFFFFF) << np.uint64(32) | np.uint64(salt & 0xFFFFFFFF)) * np.uint64(0xD6E8FEB86659FD93)
    h = _splitmix64(key)
    u = (h >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)
    out = (lo + (hi - lo) * u).astype(dtype)
    return out.reshape([s for s in shape][3 - nd:]) if nd < 3 else out
