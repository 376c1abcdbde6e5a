"""ctypes binding of the C ABI (include/yask_b200.h) -- the same entry points a cgo/JNI/C++
binding of the reference's yk_* API would use.  No torch types cross this boundary.

The library is never replaced by a CPU path: if libyask_b200.so is missing this module raises,
and on a machine without a CUDA device `Solution.prepare()` raises YaskError.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YASK_B200_LIB", os.path.join(_HERE, "lib", "libyask_b200.so"))   # env override: tuning experiments only

YB_MAX_DIMS = 5
YB_NAME_LEN = 64

FP_STRICT, FP_FMA, FP_REF_GCC = 0, 1, 2


class YaskError(RuntimeError):
    """Mirror of yask::yask_exception (/root/reference/include/yask_common_api.hpp:125-179)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"YASK error: {msg} (code {code})")
        self.code = code


class DimInfo(C.Structure):
    _fields_ = [("name", C.c_char * YB_NAME_LEN), ("kind", C.c_int32), ("domain_index", C.c_int32),
                ("rank_offset", C.c_int64), ("domain_size", C.c_int64), ("left_halo", C.c_int64), ("right_halo", C.c_int64),
                ("left_pad", C.c_int64), ("right_pad", C.c_int64), ("alloc_size", C.c_int64), ("first_misc_index", C.c_int64),
                ("stride", C.c_int64)]


class VarInfo(C.Structure):
    _fields_ = [("name", C.c_char * YB_NAME_LEN), ("num_dims", C.c_int32), ("elem_bytes", C.c_int32), ("has_step", C.c_int32),
                ("step_alloc", C.c_int32), ("first_valid_step", C.c_int64), ("last_valid_step", C.c_int64),
                ("is_output", C.c_int32), ("halo_exchange_l1_norm", C.c_int32), ("slot_elems", C.c_int64),
                ("storage_bytes", C.c_int64), ("dims", DimInfo * YB_MAX_DIMS)]


class Stats(C.Structure):
    _fields_ = [("num_elements", C.c_int64), ("num_steps_done", C.c_int64), ("num_writes_done", C.c_int64),
                ("est_fp_ops_done", C.c_int64), ("num_reads_done", C.c_int64), ("elapsed_secs", C.c_double),
                ("halo_secs", C.c_double), ("kernel_launches", C.c_int64)]


# every symbol include/yask_b200.h declares (tests check the .so exports exactly these)
ABI_SYMBOLS = [
    "yb_version_string", "yb_last_error", "yb_device_count", "yb_num_stencils", "yb_stencil_name",
    "yb_solution_create", "yb_solution_destroy", "yb_solution_name", "yb_solution_target", "yb_solution_elem_bytes",
    "yb_solution_num_domain_dims", "yb_solution_domain_dim_name", "yb_solution_step_dim_name",
    "yb_set_rank_domain_size", "yb_set_overall_domain_size", "yb_set_num_ranks", "yb_set_rank_index", "yb_set_min_pad_size",
    "yb_get_rank_domain_size", "yb_get_overall_domain_size", "yb_get_num_ranks", "yb_get_rank_index",
    "yb_get_first_rank_domain_index", "yb_get_last_rank_domain_index", "yb_set_option", "yb_get_option", "yb_set_stream",
    "yb_solution_plan_geometry", "yb_solution_prepare", "yb_solution_is_prepared", "yb_num_vars", "yb_var_index", "yb_var_info_get", "yb_var_create", "yb_var_fuse", "yb_var_set_min_pad",
    "yb_var_set_slice", "yb_var_get_slice", "yb_var_set_slice_device", "yb_var_get_slice_device", "yb_var_set_all_same",
    "yb_var_set_slice_same", "yb_var_reduce_slice", "yb_solution_auto_tune", "yb_solution_reset_auto_tuner", "yb_solution_is_auto_tuner_enabled", "yb_solution_auto_tuner_report", "yb_var_fill_hash", "yb_var_fill_hash_shifted", "yb_var_checksum", "yb_var_device_ptr", "yb_copy_to_host", "yb_solution_run", "yb_solution_sync",
    "yb_get_stats", "yb_clear_stats", "yb_halo_export_size", "yb_halo_export", "yb_halo_import", "yb_halo_finalize",
    "yb_exchange_halos", "yb_comm_env_rank", "yb_comm_env_world", "yb_comm_env_local_rank", "yb_comm_init", "yb_comm_rank", "yb_comm_world",
    "yb_comm_barrier", "yb_comm_allgather", "yb_comm_sum_i64", "yb_comm_max_f64", "yb_comm_finalize", "yb_halo_connect",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        p, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        L.yb_version_string.restype = C.c_char_p
        L.yb_last_error.restype = C.c_char_p
        L.yb_stencil_name.restype = C.c_char_p
        L.yb_stencil_name.argtypes = [i32]
        L.yb_solution_create.argtypes = [C.POINTER(p), C.c_char_p, i32, i32]
        L.yb_solution_destroy.argtypes = [p]
        for nm in ("yb_solution_name", "yb_solution_target", "yb_solution_step_dim_name"):
            getattr(L, nm).restype = C.c_char_p
            getattr(L, nm).argtypes = [p]
        L.yb_solution_domain_dim_name.restype = C.c_char_p
        L.yb_solution_domain_dim_name.argtypes = [p, i32]
        L.yb_solution_elem_bytes.argtypes = [p]
        L.yb_solution_num_domain_dims.argtypes = [p]
        for nm in ("yb_set_rank_domain_size", "yb_set_overall_domain_size", "yb_set_num_ranks", "yb_set_rank_index",
                   "yb_set_min_pad_size"):
            getattr(L, nm).argtypes = [p, i32, i64]
        for nm in ("yb_get_rank_domain_size", "yb_get_overall_domain_size", "yb_get_num_ranks", "yb_get_rank_index",
                   "yb_get_first_rank_domain_index", "yb_get_last_rank_domain_index"):
            getattr(L, nm).argtypes = [p, i32]
            getattr(L, nm).restype = i64
        L.yb_set_option.argtypes = [p, C.c_char_p, C.c_char_p]
        L.yb_get_option.argtypes = [p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.yb_solution_auto_tune.argtypes = [p, C.c_char_p, C.c_size_t]
        L.yb_solution_reset_auto_tuner.argtypes = [p, i32]
        L.yb_solution_is_auto_tuner_enabled.argtypes = [p]
        L.yb_solution_auto_tuner_report.argtypes = [p, C.c_char_p, C.c_size_t]
        L.yb_set_stream.argtypes = [p, p]
        L.yb_solution_plan_geometry.argtypes = [p]
        L.yb_solution_prepare.argtypes = [p, i32]
        L.yb_solution_is_prepared.argtypes = [p]
        L.yb_num_vars.argtypes = [p]
        L.yb_var_index.argtypes = [p, C.c_char_p]
        L.yb_var_info_get.argtypes = [p, i32, C.POINTER(VarInfo)]
        L.yb_var_create.argtypes = [p, C.c_char_p, i32, C.POINTER(C.c_char_p), C.POINTER(i64)]
        L.yb_var_set_min_pad.argtypes = [p, i32, i32, i64, i64]
        L.yb_var_fuse.argtypes = [p, i32, p, i32]
        for nm in ("yb_var_set_slice", "yb_var_get_slice", "yb_var_set_slice_device", "yb_var_get_slice_device"):
            getattr(L, nm).argtypes = [p, i32, p, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        L.yb_var_set_all_same.argtypes = [p, i32, C.c_double]
        L.yb_var_set_slice_same.argtypes = [p, i32, C.c_double, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        L.yb_var_reduce_slice.argtypes = [p, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_double), C.POINTER(i64)]
        L.yb_var_fill_hash.argtypes = [p, i32, i64, C.c_uint32, C.c_uint32, C.c_double, C.c_double]
        L.yb_var_fill_hash_shifted.argtypes = [p, i32, i64, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.POINTER(i64)]
        L.yb_var_checksum.argtypes = [p, i32, i64, C.POINTER(C.c_uint64)]
        L.yb_var_device_ptr.argtypes = [p, i32, i64, C.POINTER(p)]
        L.yb_solution_run.argtypes = [p, i64, i64]
        L.yb_solution_sync.argtypes = [p]
        L.yb_get_stats.argtypes = [p, C.POINTER(Stats)]
        L.yb_clear_stats.argtypes = [p]
        L.yb_halo_export_size.argtypes = [p, C.POINTER(C.c_size_t)]
        L.yb_halo_export.argtypes = [p, p, C.c_size_t]
        L.yb_halo_import.argtypes = [p, i64, p, C.c_size_t]
        L.yb_halo_finalize.argtypes = [p]
        L.yb_exchange_halos.argtypes = [p]
        L.yb_comm_init.argtypes = [i32, i32, C.c_char_p]
        L.yb_comm_allgather.argtypes = [p, C.c_size_t, p]
        L.yb_comm_sum_i64.argtypes = [i64, C.POINTER(i64)]
        L.yb_comm_max_f64.argtypes = [C.c_double, C.POINTER(C.c_double)]
        L.yb_halo_connect.argtypes = [p]
        _lib = L
    return _lib


def _chk(rc: int) -> int:
    if rc < 0:
        raise YaskError(rc, lib().yb_last_error().decode())
    return rc


def device_count() -> int:
    return lib().yb_device_count()


def version() -> str:
    return lib().yb_version_string().decode()


def _arr(vals: Sequence[int]):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


class Var:
    """Handle to one var of a Solution; method names follow yk_var (aux/yk_var_api.hpp)."""

    def __init__(self, soln: "Solution", index: int):
        self.soln, self.index = soln, index

    @property
    def info(self) -> VarInfo:
        vi = VarInfo()
        _chk(lib().yb_var_info_get(self.soln._h, self.index, C.byref(vi)))
        return vi

    def get_name(self) -> str:
        return self.info.name.decode()

    def get_dim_names(self):
        vi = self.info
        return [vi.dims[i].name.decode() for i in range(vi.num_dims)]

    @property
    def dtype(self):
        return np.float32 if self.info.elem_bytes == 4 else np.float64

    def _box(self, step, with_halo: bool):
        vi = self.info
        first, last = [], []
        for i in range(vi.num_dims):
            d = vi.dims[i]
            if d.kind == 0:
                first.append(step)
                last.append(step)
            elif d.kind == 2:
                first.append(d.first_misc_index)
                last.append(d.first_misc_index + d.domain_size - 1)
            else:
                first.append(d.rank_offset - (d.left_halo if with_halo else 0))
                last.append(d.rank_offset + d.domain_size - 1 + (d.right_halo if with_halo else 0))
        return first, last

    def halo_box(self, step=0):
        """(first, last) global indices of the rank's domain+halo box at API step `step`."""
        return self._box(step, True)

    def domain_box(self, step=0):
        return self._box(step, False)

    def set_elements_in_slice(self, buf: np.ndarray, first, last) -> int:
        a = np.ascontiguousarray(buf, dtype=self.dtype)
        n = C.c_int64(0)
        _chk(lib().yb_var_set_slice(self.soln._h, self.index, a.ctypes.data, _arr(first), _arr(last), C.byref(n)))
        assert n.value == a.size, (n.value, a.size)
        return n.value

    def get_elements_in_slice(self, first, last) -> np.ndarray:
        vi = self.info
        shape = [int(l - f + 1) for f, l in zip(first, last)]
        out = np.empty(shape, dtype=self.dtype)
        n = C.c_int64(0)
        _chk(lib().yb_var_get_slice(self.soln._h, self.index, out.ctypes.data, _arr(first), _arr(last), C.byref(n)))
        if vi.has_step and shape[0] == 1:
            out = out.reshape(shape[1:])
        return out

    def set_all_elements_same(self, val: float):
        _chk(lib().yb_var_set_all_same(self.soln._h, self.index, float(val)))

    def set_elements_in_slice_same(self, val: float, first, last) -> int:
        n = C.c_int64(0)
        _chk(lib().yb_var_set_slice_same(self.soln._h, self.index, float(val), _arr(first), _arr(last), C.byref(n)))
        return n.value

    def reduce_elements_in_slice(self, first, last) -> dict:
        """yk_var::reduce_elements_in_slice with every reduction requested: computed on the device in double."""
        out = (C.c_double * 5)()
        n = C.c_int64(0)
        _chk(lib().yb_var_reduce_slice(self.soln._h, self.index, _arr(first), _arr(last), out, C.byref(n)))
        return {"num": n.value, "sum": out[0], "sum_squares": out[1], "product": out[2], "max": out[3], "min": out[4]}

    def fill_hash(self, step: int, seed: int, salt: int, lo: float, hi: float, shift: Sequence[int] | None = None):
        """Hash field over the rank's halo box of `step`; `shift` (per solution domain dim) moves the window of global
        indices the values are drawn from."""
        if shift is None:
            _chk(lib().yb_var_fill_hash(self.soln._h, self.index, int(step), seed & 0xFFFFFFFF, salt & 0xFFFFFFFF, lo, hi))
        else:
            _chk(lib().yb_var_fill_hash_shifted(self.soln._h, self.index, int(step), seed & 0xFFFFFFFF, salt & 0xFFFFFFFF, lo, hi,
                                                _arr(list(shift) + [0] * (3 - len(shift)))))

    def fuse_vars(self, source: "Var"):
        """yk_var::fuse_vars: this var becomes another reference to `source`'s device storage."""
        _chk(lib().yb_var_fuse(self.soln._h, self.index, source.soln._h, source.index))

    def checksum(self, step: int) -> int:
        out = C.c_uint64(0)
        _chk(lib().yb_var_checksum(self.soln._h, self.index, int(step), C.byref(out)))
        return out.value

    def device_ptr(self, step: int = 0) -> int:
        out = C.c_void_p()
        _chk(lib().yb_var_device_ptr(self.soln._h, self.index, int(step), C.byref(out)))
        return out.value

    def get_first_valid_step_index(self):
        return self.info.first_valid_step

    def get_last_valid_step_index(self):
        return self.info.last_valid_step


class Solution:
    """Mirror of yk_solution for the hot path (aux/yk_solution_api.hpp): sizes, prepare, vars, run, stats."""

    def __init__(self, stencil: str = "iso3dfd", radius: int = 0, elem_bytes: int = 4):
        h = C.c_void_p()
        _chk(lib().yb_solution_create(C.byref(h), stencil.encode(), radius, elem_bytes))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().yb_solution_destroy(self._h)
            self._h = None

    end_solution = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_name(self):
        return lib().yb_solution_name(self._h).decode()

    def get_target(self):
        return lib().yb_solution_target(self._h).decode()

    def get_element_bytes(self):
        return lib().yb_solution_elem_bytes(self._h)

    def get_domain_dim_names(self):
        return [lib().yb_solution_domain_dim_name(self._h, i).decode() for i in range(lib().yb_solution_num_domain_dims(self._h))]

    def get_step_dim_name(self):
        return lib().yb_solution_step_dim_name(self._h).decode()

    def _dim(self, dim):
        return dim if isinstance(dim, int) else self.get_domain_dim_names().index(dim)

    def set_overall_domain_size_vec(self, n):
        for d, v in enumerate(n):
            _chk(lib().yb_set_overall_domain_size(self._h, d, int(v)))

    def set_rank_domain_size_vec(self, n):
        for d, v in enumerate(n):
            _chk(lib().yb_set_rank_domain_size(self._h, d, int(v)))

    def set_num_ranks_vec(self, n):
        for d, v in enumerate(n):
            _chk(lib().yb_set_num_ranks(self._h, d, int(v)))

    def set_rank_index_vec(self, n):
        for d, v in enumerate(n):
            _chk(lib().yb_set_rank_index(self._h, d, int(v)))

    def get_rank_domain_size_vec(self):
        return [lib().yb_get_rank_domain_size(self._h, d) for d in range(len(self.get_domain_dim_names()))]

    def get_overall_domain_size_vec(self):
        return [lib().yb_get_overall_domain_size(self._h, d) for d in range(len(self.get_domain_dim_names()))]

    def get_first_rank_domain_index_vec(self):
        return [lib().yb_get_first_rank_domain_index(self._h, d) for d in range(len(self.get_domain_dim_names()))]

    def set_option(self, key: str, value):
        _chk(lib().yb_set_option(self._h, key.encode(), str(value).encode()))

    def run_auto_tuner_now(self) -> str:
        """Offline auto-tuner: times the engine's launch variants, keeps the fastest; var contents are not preserved."""
        buf = C.create_string_buffer(8192)
        _chk(lib().yb_solution_auto_tune(self._h, buf, 8192))
        return buf.value.decode()

    def reset_auto_tuner(self, enable: bool = True):
        _chk(lib().yb_solution_reset_auto_tuner(self._h, 1 if enable else 0))

    def is_auto_tuner_enabled(self) -> bool:
        return bool(lib().yb_solution_is_auto_tuner_enabled(self._h))

    def auto_tuner_report(self) -> str:
        buf = C.create_string_buffer(8192)
        _chk(lib().yb_solution_auto_tuner_report(self._h, buf, 8192))
        return buf.value.decode()

    def get_option(self, key: str) -> str:
        buf = C.create_string_buffer(256)
        _chk(lib().yb_get_option(self._h, key.encode(), buf, 256))
        return buf.value.decode()

    def set_stream(self, cuda_stream: int):
        _chk(lib().yb_set_stream(self._h, C.c_void_p(cuda_stream)))

    def plan_geometry(self):
        _chk(lib().yb_solution_plan_geometry(self._h))

    def prepare_solution(self, device: int = 0):
        _chk(lib().yb_solution_prepare(self._h, device))

    def get_num_vars(self):
        return lib().yb_num_vars(self._h)

    def get_var(self, name: str) -> Var:
        return Var(self, _chk(lib().yb_var_index(self._h, name.encode())))

    def new_fixed_size_var(self, name: str, dims, sizes) -> Var:
        names = (C.c_char_p * len(dims))(*[d.encode() for d in dims])
        return Var(self, _chk(lib().yb_var_create(self._h, name.encode(), len(dims), names, _arr(sizes))))

    def new_var(self, name: str, dims) -> Var:
        names = (C.c_char_p * len(dims))(*[d.encode() for d in dims])
        return Var(self, _chk(lib().yb_var_create(self._h, name.encode(), len(dims), names, None)))

    def get_vars(self):
        return [Var(self, i) for i in range(self.get_num_vars())]

    def run_solution(self, first_step: int, last_step: int | None = None):
        _chk(lib().yb_solution_run(self._h, first_step, first_step if last_step is None else last_step))

    def sync(self):
        _chk(lib().yb_solution_sync(self._h))

    def get_stats(self) -> Stats:
        st = Stats()
        _chk(lib().yb_get_stats(self._h, C.byref(st)))
        return st

    def clear_stats(self):
        _chk(lib().yb_clear_stats(self._h))

    # multi-GPU plumbing (yb_halo_*)
    def halo_export(self) -> bytes:
        n = C.c_size_t(0)
        _chk(lib().yb_halo_export_size(self._h, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _chk(lib().yb_halo_export(self._h, buf, n.value))
        return buf.raw

    def halo_import(self, peer_rank: int, blob: bytes):
        _chk(lib().yb_halo_import(self._h, peer_rank, blob, len(blob)))

    def halo_finalize(self):
        _chk(lib().yb_halo_finalize(self._h))

    def exchange_halos(self):
        _chk(lib().yb_exchange_halos(self._h))
