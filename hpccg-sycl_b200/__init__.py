"""hpccg-sycl_b200 -- Python face of the B200-native HPCCG hot path (tests, bench.py, launch plumbing).

The product is libhpccg_b200.so (hand-written sm_100a CUDA kernels behind the C-ABI of
include/hpccg_b200.h plus the reference-named C++ API).  This package only binds it with ctypes and
mirrors the reference's operator interface with the same names and argument meaning:

    generate_matrix(nx, ny, nz)            generate_matrix.hpp:58
    make_local_matrix(A)                   make_local_matrix.hpp:48
    HPCCG(A, b, x, max_iter, tolerance)    HPCCG.hpp:61-63
    HPC_sparsemv(A, x, y)                  HPC_sparsemv.hpp:55-56
    ddot(n, x, y)                          ddot.hpp:55-56
    waxpby(n, alpha, x, beta, y, w)        waxpby.hpp:51-53
    exchange_externals(A, x)               exchange_externals.hpp:49
    compute_residual(n, v1, v2)            compute_residual.hpp:50-51

Vectors may be numpy arrays (host pointers, as the reference's callers pass), torch CUDA tensors or
raw integer addresses.  PyTorch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Callable, Sequence

import numpy as np

from . import _capi
from ._capi import HpccgError, check, lib

__all__ = [
    "HpccgError", "Matrix", "DeviceMatrix", "set_matrix_format", "set_rank", "get_rank", "set_options", "set_print", "generate_matrix", "read_HPC_row",
    "read_HPC_row", "make_local_matrix", "HPCCG", "HPC_sparsemv", "ddot", "waxpby", "exchange_externals", "compute_residual",
    "yaml_report", "run_local_world", "launch_count", "dev",
]

SOLVE_DEFAULT, SOLVE_UNFUSED, SOLVE_NO_OVERLAP, SOLVE_TIMERS, SOLVE_NCCL_ONLY, SOLVE_GRAPH, SOLVE_EAGER_X, SOLVE_PERSISTENT = 0, 1, 2, 4, 8, 16, 32, 64


def _ptr(v) -> int:
    """Address of a numpy array, torch tensor, ctypes pointer or int."""
    if v is None:
        return 0
    if isinstance(v, int):
        return v
    if isinstance(v, np.ndarray):
        if not v.flags.c_contiguous:
            raise ValueError("array must be contiguous")
        return v.ctypes.data
    if hasattr(v, "data_ptr"):
        if not v.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return v.data_ptr()
    if isinstance(v, C._Pointer) or isinstance(v, C.c_void_p):
        return C.cast(v, C.c_void_p).value or 0
    raise TypeError(f"cannot take the address of {type(v)}")


def _stream(stream) -> int:
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)  # torch.cuda.Stream


def launch_count() -> int:
    return int(lib.hpccg_launch_count())


# ---- rank context -----------------------------------------------------------------------------------------------
def set_rank(rank: int, size: int) -> None:
    check(lib.hpccg_ctx_set(rank, size), "hpccg_ctx_set")


def get_rank() -> tuple[int, int]:
    r, s = C.c_int(), C.c_int()
    check(lib.hpccg_ctx_get(C.byref(r), C.byref(s)))
    return r.value, s.value


def set_options(stencil: int = 27, host_arrays: bool = True) -> None:
    """The two switches the reference fixes at compile time (generate_matrix.cpp:219; host staging)."""
    check(lib.hpccg_api_set_options(stencil, 1 if host_arrays else 0), "hpccg_api_set_options")


_ARRAY_DTYPES = {
    "nnz_in_row": np.int32, "list_of_inds": np.int32, "list_of_vals": np.float64, "ind_offsets": np.int64,
    "val_offsets": np.int64, "diag_offsets": np.int64, "external_index": np.int32, "external_local_index": np.int32,
    "elements_to_send": np.int32, "neighbors": np.int32, "recv_length": np.int32, "send_length": np.int32,
}


class DeviceMatrix:
    """Opaque hpccg_dev_matrix* (column-major ELLPACK in HBM)."""

    def __init__(self, handle: int, owned: bool):
        self.handle = handle
        self.owned = owned

    @classmethod
    def generate(cls, nx, ny, nz, rank=0, size=1, stencil=27, lower=None, upper=None, local_ncol=None):
        out = C.c_void_p()
        n = nx * ny * nz
        if local_ncol is None:
            local_ncol = n + (nx * ny if lower is not None else 0) + (nx * ny if upper is not None else 0)
        check(lib.hpccg_dev_matrix_generate(nx, ny, nz, rank, size, stencil, _ptr(lower), _ptr(upper), local_ncol,
                                            C.byref(out)), "hpccg_dev_matrix_generate")
        return cls(out.value, True)

    def info(self):
        n, nc, s, npad = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        check(lib.hpccg_dev_matrix_info(self.handle, C.byref(n), C.byref(nc), C.byref(s), C.byref(npad)))
        return {"local_nrow": n.value, "local_ncol": nc.value, "slots": s.value, "padded_rows": npad.value}

    def download(self):
        i = self.info()
        vals = np.empty((i["slots"], i["padded_rows"]), dtype=np.float64)
        cols = np.empty((i["slots"], i["padded_rows"]), dtype=np.int32)
        check(lib.hpccg_dev_matrix_download(self.handle, vals.ctypes.data, cols.ctypes.data))
        return vals, cols

    def compress(self) -> dict:
        """Pattern-coded re-encoding (lossless, bit-identical SpMV); returns format()."""
        check(lib.hpccg_dev_matrix_compress(self.handle), "hpccg_dev_matrix_compress")
        return self.format()

    def format(self) -> dict:
        f, p = C.c_int(), C.c_int()
        check(lib.hpccg_dev_matrix_format(self.handle, C.byref(f), C.byref(p)))
        return {"format": f.value, "patterns": p.value}

    def comm(self) -> dict:
        """Data plane of the multi-rank solves on this mirror: peer memory inside the kernels or NCCL between them."""
        p, f = C.c_int(), C.c_int()
        check(lib.hpccg_dev_matrix_comm(self.handle, C.byref(p), C.byref(f)))
        return {"peer": bool(p.value), "fused_put": bool(f.value)}

    def bytes(self) -> int:
        b = C.c_longlong()
        check(lib.hpccg_dev_matrix_bytes(self.handle, C.byref(b)))
        return b.value

    def destroy(self):
        if self.owned and self.handle:
            lib.hpccg_dev_matrix_destroy(self.handle)
        self.handle = None


class Matrix:
    """HPC_Sparse_Matrix* built by generate_matrix (HPC_Sparse_Matrix.hpp:54-85 field names)."""

    def __init__(self, handle: int, x: np.ndarray, b: np.ndarray, xexact: np.ndarray, raw):
        self.handle = handle
        self.x, self.b, self.xexact = x, b, xexact
        self._raw = raw

    def scalar(self, name: str) -> int:
        return int(lib.hpccg_api_matrix_scalar(self.handle, name.encode()))

    def array(self, name: str) -> np.ndarray:
        n = lib.hpccg_api_matrix_array(self.handle, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=_ARRAY_DTYPES[name])
        lib.hpccg_api_matrix_array(self.handle, name.encode(), out.ctypes.data, n)
        return out

    def __getattr__(self, name):
        if name in ("start_row", "stop_row", "total_nrow", "total_nnz", "local_nrow", "local_ncol", "local_nnz",
                    "num_external", "num_send_neighbors", "total_to_be_sent"):
            return self.scalar(name)
        raise AttributeError(name)

    def device(self) -> DeviceMatrix:
        """The ELL mirror (created now if it does not exist yet; outside any timed region)."""
        out = C.c_void_p()
        check(lib.hpccg_api_matrix_device(self.handle, C.byref(out)), "hpccg_api_matrix_device")
        return DeviceMatrix(out.value, False)

    def destroy(self):
        if self.handle:
            lib.hpccg_api_destroyMatrix(self.handle)
            lib.hpccg_api_free_vectors(*self._raw)
            self.handle = None
            self.x = self.b = self.xexact = None


def set_matrix_format(fmt) -> None:
    """Device-mirror format for matrices created from now on: 0 / "sell" (default) or 1 / "pattern"."""
    fmt = {"sell": 0, "pattern": 1}.get(fmt, fmt)
    check(lib.hpccg_api_set_matrix_format(int(fmt)), "hpccg_api_set_matrix_format")


def set_print(on: bool) -> None:
    """Residual lines of HPCCG.cpp:356,372-373 on rank 0 (default on, like the reference)."""
    check(lib.hpccg_api_set_print(1 if on else 0))


def generate_matrix(nx: int, ny: int, nz: int) -> Matrix:
    """generate_matrix.cpp:196-307; returns the matrix with its x (zeros), b (= A*1) and xexact (ones)."""
    A = C.c_void_p()
    x, b, e = _capi.PD(), _capi.PD(), _capi.PD()
    check(lib.hpccg_api_generate_matrix(nx, ny, nz, C.byref(A), C.byref(x), C.byref(b), C.byref(e)), "generate_matrix")
    n = nx * ny * nz
    views = [np.ctypeslib.as_array(p, shape=(n,)) for p in (x, b, e)]
    return Matrix(A.value, views[0], views[1], views[2], (x, b, e))


def read_HPC_row(data_file) -> Matrix:
    """read_HPC_row.cpp:217-373: the matrix, x, b and xexact of this rank from the reference's text format."""
    A = C.c_void_p()
    x, b, e = _capi.PD(), _capi.PD(), _capi.PD()
    check(lib.hpccg_api_read_HPC_row(str(data_file).encode(), C.byref(A), C.byref(x), C.byref(b), C.byref(e)), "read_HPC_row")
    n = int(lib.hpccg_api_matrix_scalar(A.value, b"local_nrow"))
    views = [np.ctypeslib.as_array(p, shape=(n,)) for p in (x, b, e)]
    return Matrix(A.value, views[0], views[1], views[2], (x, b, e))


def make_local_matrix(A: Matrix) -> None:
    check(lib.hpccg_api_make_local_matrix(A.handle), "make_local_matrix")


def HPCCG(A: Matrix, b, x, max_iter: int = 150, tolerance: float = 0.0):
    """HPCCG.cpp:312-402.  Returns (niters, normr, times[7], history[max_iter])."""
    niters, normr = C.c_int(), C.c_double()
    times = np.zeros(7)
    check(lib.hpccg_api_HPCCG(A.handle, _ptr(b), _ptr(x), max_iter, tolerance, C.byref(niters), C.byref(normr),
                              times.ctypes.data_as(_capi.PD)), "HPCCG")
    hist = np.full(max(max_iter, 1), np.nan)
    lib.hpccg_api_last_history(hist.ctypes.data_as(_capi.PD), len(hist))
    return niters.value, normr.value, times, hist


def HPC_sparsemv(A: Matrix, x, y) -> None:
    check(lib.hpccg_api_HPC_sparsemv(A.handle, _ptr(x), _ptr(y)), "HPC_sparsemv")


def ddot(n: int, x, y):
    """Returns (result, time_allreduce)."""
    r, t = C.c_double(), C.c_double()
    check(lib.hpccg_api_ddot(n, _ptr(x), _ptr(y), C.byref(r), C.byref(t)), "ddot")
    return r.value, t.value


def waxpby(n: int, alpha: float, x, beta: float, y, w) -> None:
    check(lib.hpccg_api_waxpby(n, alpha, _ptr(x), beta, _ptr(y), _ptr(w)), "waxpby")


def exchange_externals(A: Matrix, x) -> None:
    check(lib.hpccg_api_exchange_externals(A.handle, _ptr(x)), "exchange_externals")


def compute_residual(n: int, v1, v2) -> float:
    r = C.c_double()
    check(lib.hpccg_api_compute_residual(n, _ptr(v1), _ptr(v2), C.byref(r)), "compute_residual")
    return r.value


def yaml_report(nx, ny, nz, niters, normr, times, total_nrow, total_nnz, ranks=0, omp_threads=0, t4stats=None) -> str:
    times = np.ascontiguousarray(times, dtype=np.float64)
    t4 = np.ascontiguousarray(t4stats if t4stats is not None else [0.0, 0.0, 0.0], dtype=np.float64)
    buf = C.create_string_buffer(1 << 16)
    n = lib.hpccg_api_yaml_report(nx, ny, nz, niters, normr, times.ctypes.data_as(_capi.PD), float(total_nrow),
                                  float(total_nnz), ranks, omp_threads, t4.ctypes.data_as(_capi.PD), buf, len(buf))
    if n < 0:
        raise HpccgError("yaml buffer too small")
    return buf.value.decode()


def run_local_world(size: int, fn: Callable[[int], object]) -> list:
    """Runs fn(rank) on `size` host threads bound to an in-process world (set-up collectives work)."""
    world = C.c_void_p()
    check(lib.hpccg_local_world_create(size, C.byref(world)))
    results: list = [None] * size
    errors: list = [None] * size

    def body(r):
        try:
            check(lib.hpccg_local_world_bind(world, r))
            results[r] = fn(r)
        except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
            errors[r] = e

    threads = [threading.Thread(target=body, args=(r,)) for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    lib.hpccg_local_world_destroy(world)
    for e in errors:
        if e is not None:
            raise e
    return results


class _Dev:
    """hpccg_dev_* : device pointers, asynchronous on a stream."""

    @staticmethod
    def spmv(m: DeviceMatrix, x, y, stream=None):
        check(lib.hpccg_dev_spmv(m.handle, _ptr(x), _ptr(y), _stream(stream)), "hpccg_dev_spmv")

    @staticmethod
    def spmv_dot(m: DeviceMatrix, x, y, result, stream=None):
        check(lib.hpccg_dev_spmv_dot(m.handle, _ptr(x), _ptr(y), _ptr(result), _stream(stream)), "hpccg_dev_spmv_dot")

    @staticmethod
    def dot(n, x, y, result, stream=None):
        check(lib.hpccg_dev_dot(n, _ptr(x), _ptr(y), _ptr(result), _stream(stream)), "hpccg_dev_dot")

    @staticmethod
    def waxpby(n, alpha, x, beta, y, w, stream=None):
        check(lib.hpccg_dev_waxpby(n, alpha, _ptr(x), beta, _ptr(y), _ptr(w), _stream(stream)), "hpccg_dev_waxpby")

    @staticmethod
    def update_xr_dot(n, alpha_dev, p, Ap, x, r, rr, stream=None):
        check(lib.hpccg_dev_update_xr_dot(n, _ptr(alpha_dev), _ptr(p), _ptr(Ap), _ptr(x), _ptr(r), _ptr(rr),
                                          _stream(stream)), "hpccg_dev_update_xr_dot")

    @staticmethod
    def p_update(n, beta_dev, r, p, stream=None):
        check(lib.hpccg_dev_p_update(n, _ptr(beta_dev), _ptr(r), _ptr(p), _stream(stream)), "hpccg_dev_p_update")

    @staticmethod
    def halo_pack(m: DeviceMatrix, x, send_buffer, stream=None):
        check(lib.hpccg_dev_halo_pack(m.handle, _ptr(x), _ptr(send_buffer), _stream(stream)), "hpccg_dev_halo_pack")

    @staticmethod
    def max_abs_diff(n, v1, v2, result, stream=None):
        check(lib.hpccg_dev_max_abs_diff(n, _ptr(v1), _ptr(v2), _ptr(result), _stream(stream)), "hpccg_dev_max_abs_diff")

    @staticmethod
    def cg_solve(m: DeviceMatrix, b, x, max_iter=150, tolerance=0.0, flags=0, stream=None, want_hist=True,
                 want_times=False):
        niters, normr, loop_ms = C.c_int(), C.c_double(), C.c_double()
        hist = np.full(max(max_iter, 1), np.nan) if want_hist else None
        times = np.zeros(16) if want_times else None
        check(lib.hpccg_dev_cg_solve(m.handle, _ptr(b), _ptr(x), max_iter, tolerance, C.byref(niters), C.byref(normr),
                                     _ptr(hist), _ptr(times), C.byref(loop_ms), flags, _stream(stream)),
              "hpccg_dev_cg_solve")
        return {"niters": niters.value, "normr": normr.value, "hist": hist, "times": times, "loop_ms": loop_ms.value}

    @staticmethod
    def cg_solve_group(ms: Sequence[DeviceMatrix], bs, xs, max_iter=150, tolerance=0.0, flags=0, stream=None):
        n = len(ms)
        arr_m = (C.c_void_p * n)(*[m.handle for m in ms])
        arr_b = (C.c_void_p * n)(*[_ptr(b) for b in bs])
        arr_x = (C.c_void_p * n)(*[_ptr(x) for x in xs])
        niters, normr, loop_ms = C.c_int(), C.c_double(), C.c_double()
        hist = np.full(max(max_iter, 1), np.nan)
        check(lib.hpccg_dev_cg_solve_group(n, arr_m, arr_b, arr_x, max_iter, tolerance, C.byref(niters), C.byref(normr),
                                           _ptr(hist), C.byref(loop_ms), flags, _stream(stream)),
              "hpccg_dev_cg_solve_group")
        return {"niters": niters.value, "normr": normr.value, "hist": hist, "loop_ms": loop_ms.value}


dev = _Dev()
