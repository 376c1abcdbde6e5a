"""TEST INFRASTRUCTURE ONLY: ctypes wrapper around oracle/libyask_oracle.so and a runner for
the prebuilt reference binaries under oracle/_ref (built by oracle/build_ref.sh).

Never imported by the product package (yask_b200)."""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile the C restatement (gcc, seconds)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True, env={**os.environ, "CC": "gcc"})
    return os.path.join(HERE, "libyask_oracle.so")


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libyask_oracle.so")
        src = os.path.join(HERE, "yask_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.yo_iso3dfd_coeffs.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _LIB.yo_const_roundtrip.argtypes = [ctypes.c_double]
        _LIB.yo_const_roundtrip.restype = ctypes.c_double
        for nm in ("yo_iso3dfd_run_f32", "yo_iso3dfd_run_f64"):
            f = getattr(_LIB, nm)
            f.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64] * 4 + [ctypes.c_int] * 3
            f.restype = ctypes.c_int
    return _LIB


def iso3dfd_coeffs(radius: int) -> np.ndarray:
    c = np.zeros(radius + 1, dtype=np.float64)
    rc = lib().yo_iso3dfd_coeffs(c.ctypes.data, radius)
    assert rc == 0
    return c


def iso3dfd_run(p0: np.ndarray, p1: np.ndarray, v: np.ndarray, radius: int, steps: int, contract: bool) -> np.ndarray:
    """p0/p1: (nx+2h, ny+2h, nz+2h) API steps 0 and 1 incl. halo h; v: (nx,ny,nz).
    Returns the final p (domain + untouched halo) after `steps` steps; inputs are not modified."""
    assert p0.dtype == p1.dtype == v.dtype and p0.shape == p1.shape
    nx, ny, nz = v.shape
    h = (p0.shape[0] - nx) // 2
    assert h >= radius and p0.shape == (nx + 2 * h, ny + 2 * h, nz + 2 * h)
    a = np.ascontiguousarray(p0).copy()
    b = np.ascontiguousarray(p1).copy()
    vv = np.ascontiguousarray(v)
    fn = lib().yo_iso3dfd_run_f32 if p0.dtype == np.float32 else lib().yo_iso3dfd_run_f64
    which = fn(a.ctypes.data, b.ctypes.data, vv.ctypes.data, nx, ny, nz, h, radius, steps, int(contract))
    assert which in (0, 1)
    return a if which == 0 else b


# ---------------------------------------------------------------------------------------
# Emitter-generated solutions (oracle/gen/*.gen.h): generic runner
# ---------------------------------------------------------------------------------------
class _GenArgs(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int64), ("ny", ctypes.c_int64), ("nz", ctypes.c_int64), ("ptr", ctypes.c_void_p * 96),
                ("sx", ctypes.c_int64 * 96), ("sy", ctypes.c_int64 * 96), ("sz", ctypes.c_int64 * 96),
                ("off", ctypes.c_int64 * 3), ("gfirst", ctypes.c_int64 * 3), ("glast", ctypes.c_int64 * 3),
                ("bx", ctypes.c_int64), ("by", ctypes.c_int64), ("bz", ctypes.c_int64), ("t", ctypes.c_int64)]


def gen_ir(stencil: str) -> dict:
    import json
    return json.load(open(os.path.join(os.path.dirname(HERE), "yask_b200", "csrc", "gen", f"{stencil}.json")))


def gen_run(stencil: str, n, steps: int, inputs: dict, contract: int = 0) -> dict:
    """Run `steps` steps of a generated solution on the CPU.
    contract = 0: every operation rounded separately (== the reference built with -ffp-contract=off);
    contract = 1: products fused into additions exactly where GCC fuses them in the reference's default build.
    inputs: {(var, api_step): ndarray over the var's rank halo box} (scalars: 0-d / size-1 arrays); API steps
    0..alloc_t-1 for vars with a step dim.  Returns {var: (last_valid_step, ndarray over the halo box)} for every
    written var.  Inputs are not modified."""
    ir = gen_ir(stencil)
    dd = ir["domain_dims"]
    dt = np.float32 if ir["elem_bytes"] == 4 else np.float64
    L = lib()
    L.yo_gen_run_part.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_GenArgs)]
    L.yo_gen_set_contract(1 if contract else 0)
    slots = {}      # var -> list of arrays per slot
    meta = {}
    nn = [1] * (3 - len(n)) + list(n)      # domain dims right-aligned into the (x,y,z) slots, like the CUDA engine
    sh = 3 - len(dd)
    for v in ir["vars"]:
        has_step = bool(v["dims"]) and v["dims"][0] == ir["step_dim"]
        a = v["alloc_t"] if has_step else 1
        if v.get("scratch"):
            # engine-internal temporary: rank domain + its halo, zero-initialised
            shape = [n[dd.index(d)] + v["halo"][d][0] + v["halo"][d][1] for d in v["dims"]]
            slots[v["name"]] = [np.zeros(shape, dtype=dt)]
            meta[v["name"]] = (v, False, 1)
            continue
        arrs = [np.array(inputs[(v["name"], t)], dtype=dt, copy=True, order="C") for t in range(a)]
        # API step t lives in slot t % alloc_t (imod_flr)
        slots[v["name"]] = [None] * a
        for t in range(a):
            slots[v["name"]][t % a] = arrs[t]
        meta[v["name"]] = (v, has_step, a)
    last = {}
    for t in range(steps):
        for st in ir["stages"]:
            written = set()      # scratch vars written so far in this stage
            for p in st["parts"]:
                A = _GenArgs()
                A.nx, A.ny, A.nz = nn
                A.t = t
                wh = [[0, 0]] * sh + (p.get("wh") or [[0, 0]] * len(dd))
                A.bx, A.by, A.bz = [-w[0] for w in wh]
                A.nx, A.ny, A.nz = [nn[k] + wh[k][1] for k in range(3)]
                for d in range(3):
                    A.off[d], A.gfirst[d], A.glast[d] = 0, 0, nn[d] - 1
                for k, acc in enumerate(p["accesses"]):
                    v, has_step, a = meta[acc["var"]]
                    arr = slots[acc["var"]][(t + acc["toff"]) % a]
                    vd = [d for d in v["dims"] if d != ir["step_dim"]]
                    strides = {d: arr.strides[i] // arr.itemsize for i, d in enumerate(vd)}
                    off = sum(v["halo"][d][0] * strides[d] for d in vd if d in v["halo"])
                    mdims = [d for d in vd if d not in v["halo"]]          # misc dims: constant index selects a sub-array
                    for d, mi in zip(mdims, acc.get("misc", [])):
                        off += (mi - v["misc_range"][d][0]) * strides[d]
                    A.ptr[k] = arr.ctypes.data + off * arr.itemsize
                    A.sx[k] = strides.get(dd[0 - sh], 0) if sh <= 0 else 0
                    A.sy[k] = strides.get(dd[1 - sh], 0) if sh <= 1 else 0
                    A.sz[k] = strides.get(dd[2 - sh], 0)
                if p.get("scratch"):
                    # a scratch var whose first writer in the stage is conditional starts from zero
                    # (/root/reference/src/kernel/lib/stencil_calc.cpp:85-109)
                    for o in p["outputs"]:
                        sv = p["accesses"][o["access"]]["var"]
                        if sv not in written:
                            written.add(sv)
                            if p.get("conditional"):
                                slots[sv][0][...] = 0
                rc = L.yo_gen_run_part(stencil.encode(), p.get("index", 0), ctypes.byref(A))
                assert rc == 0, rc
                if not p.get("scratch"):
                    for o in p["outputs"]:
                        oa = p["accesses"][o["access"]]
                        last[oa["var"]] = t + oa["toff"]      # t+1 (t-1 for reverse-time solutions)
    out = {}
    for name, tl in last.items():
        v, has_step, a = meta[name]
        out[name] = (tl, slots[name][tl % a])
    return out


# ---------------------------------------------------------------------------------------
# Prebuilt reference (oracle/_ref): only usable where build_ref.sh has been run.
# ---------------------------------------------------------------------------------------
# The build container has the reference's whole output tree (_ref/yask); the GPU box only what build_ref.sh staged (_ref/ship).
REF_BIN = os.path.join(HERE, "_ref", "yask", "bin")
if not os.path.isdir(REF_BIN):
    REF_BIN = os.path.join(HERE, "_ref", "ship", "bin")


def ref_available(tag: str) -> bool:
    return os.path.exists(os.path.join(REF_BIN, f"ref_driver.{tag}"))


def _parse_manifest(text: str) -> dict:
    out = {"vars": {}}
    for line in text.splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "solution":
            out["name"] = tok[1]
            out["elem_bytes"] = int(tok[3])
        elif tok[0] == "var":
            name = tok[1]
            nd = int(tok[3])
            i = 5
            dims = tok[i:i + nd]
            i += nd
            has_step = int(tok[i + 1])
            steps = (int(tok[i + 3]), int(tok[i + 4]))
            i += 5
            nb = nd - has_step
            boxes = {}
            for key in ("in_first", "in_last", "out_first", "out_last"):
                assert tok[i] == key, (tok[i], key)
                boxes[key] = [int(t) for t in tok[i + 1:i + 1 + nb]]
                i += 1 + nb
            out["vars"][name] = dict(dims=dims, has_step=bool(has_step), steps=steps, **boxes)
    return out


def ref_info(tag: str, n) -> dict:
    n = list(n) + [1] * (3 - len(n))     # the driver takes three sizes and uses the first ndd
    r = subprocess.run([os.path.join(REF_BIN, f"ref_driver.{tag}"), "info"] + [str(int(i)) for i in n],
                       check=True, capture_output=True, text=True)
    return _parse_manifest(r.stdout)


def ref_run(tag: str, n, steps: int, inputs: dict, threads: int | None = None) -> tuple[dict, dict]:
    """inputs: {(var, step): ndarray over the var's in-box}.  Returns ({(var, step): ndarray over
    the out-box}, manifest-after-run)."""
    n = list(n) + [1] * (3 - len(n))
    info = ref_info(tag, n)
    dt = np.float32 if info["elem_bytes"] == 4 else np.float64
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    with tempfile.TemporaryDirectory() as d:
        for name, g in info["vars"].items():
            t0, t1 = g["steps"] if g["has_step"] else (0, 0)
            shape = [l - f + 1 for f, l in zip(g["in_first"], g["in_last"])]
            for t in range(t0, t1 + 1):
                a = np.ascontiguousarray(inputs[(name, t)], dtype=dt)
                assert list(a.shape) == shape or (not shape and a.size == 1), (name, a.shape, shape)
                a.tofile(os.path.join(d, f"{name}.t{t}.in"))
        r = subprocess.run([os.path.join(REF_BIN, f"ref_driver.{tag}"), "run"] + [str(int(i)) for i in n] + [str(steps), d],
                           check=True, capture_output=True, text=True, env=env)
        after = _parse_manifest(r.stdout)
        outs = {}
        for name, g in after["vars"].items():
            t0, t1 = g["steps"] if g["has_step"] else (0, 0)
            shape = [l - f + 1 for f, l in zip(g["out_first"], g["out_last"])]
            for t in range(t0, t1 + 1):
                a = np.fromfile(os.path.join(d, f"{name}.t{t}.out"), dtype=dt)
                outs[(name, t)] = a.reshape(shape) if shape else a
    return outs, after
