"""oracle/refwrap.py -- TEST INFRASTRUCTURE, not product code.

ctypes access to oracle/_ref/libhpccg_ref_{serial,omp,mpi}.so: the REAL reference
(/root/reference sources compiled unmodified by oracle/build.sh) behind the C-ABI
of oracle/ref_driver.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_REFDIR = _HERE / "_ref"

_PD = C.POINTER(C.c_double)
_PPD = C.POINTER(_PD)


def _path(variant: str) -> Path:
    # "oracle" = the C restatement (hpccg_oracle.c); the rest = the real reference
    return _REFDIR / ("libhpccg_oracle.so" if variant == "oracle" else f"libhpccg_ref_{variant}.so")


def available(variant: str = "serial") -> bool:
    return _path(variant).exists()


class _Prefixed:
    """Maps lib.ref_foo onto orc_foo for the C restatement, which has the same C-ABI."""

    def __init__(self, cdll, prefix):
        self._cdll, self._prefix = cdll, prefix

    def __getattr__(self, name):
        if name.startswith("ref_"):
            name = self._prefix + name[4:]
        return getattr(self._cdll, name)


_libs: dict = {}


def _lib(variant: str):
    if variant in _libs:
        return _libs[variant]
    path = _path(variant)
    if not path.exists():
        raise FileNotFoundError(f"{path} missing: run oracle/build.sh (the ref_* variants need /root/reference)")
    # RTLD_LOCAL: the three reference variants export the same symbols.
    lib = _Prefixed(C.CDLL(str(path), mode=os.RTLD_LOCAL | os.RTLD_NOW), "orc_" if variant == "oracle" else "ref_")
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_int] * 5
    lib.ref_destroy.argtypes = [C.c_void_p]
    lib.ref_scalar.restype = C.c_longlong
    lib.ref_scalar.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.ref_array.restype = C.c_longlong
    lib.ref_array.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_longlong]
    lib.ref_spmv.argtypes = [C.c_void_p, _PPD, _PPD, C.c_int, C.c_int]
    lib.ref_ddot.argtypes = [C.c_void_p, _PPD, _PPD, _PD]
    lib.ref_ddot_raw.argtypes = [C.c_int, _PD, _PD, _PD]
    lib.ref_waxpby.argtypes = [C.c_int, C.c_double, _PD, C.c_double, _PD, _PD]
    lib.ref_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, _PD, C.POINTER(C.c_int), _PD, _PD, _PPD]
    lib.ref_compute_residual.argtypes = [C.c_void_p, _PPD, _PD]
    if variant != "oracle":
        lib.ref_create_from_file.restype = C.c_void_p
        lib.ref_create_from_file.argtypes = [C.c_char_p, C.c_int]
        lib.ref_yaml_report.argtypes = [C.c_int] * 4 + [C.c_double, _PD, C.c_double, C.c_double, C.c_int, C.c_int,
                                                        _PD, C.c_char_p, C.c_int]
    _libs[variant] = lib
    return lib


def _pd(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_PD)


def _ppd(arrs):
    arr = (_PD * len(arrs))(*[_pd(a) if a is not None else _PD() for a in arrs])
    return arr


_DTYPES = {
    "nnz_in_row": np.int32, "list_of_inds": np.int32, "list_of_vals": np.float64, "x": np.float64, "b": np.float64,
    "xexact": np.float64, "ind_offsets": np.int64, "val_offsets": np.int64, "diag_offsets": np.int64,
    "external_index": np.int32, "external_local_index": np.int32, "elements_to_send": np.int32,
    "neighbors": np.int32, "recv_length": np.int32, "send_length": np.int32,
}

SCALARS = ("start_row", "stop_row", "total_nrow", "total_nnz", "local_nrow", "local_ncol", "local_nnz", "nnz_sum",
           "num_external", "num_send_neighbors", "total_to_be_sent")


class RefWorld:
    """The reference's matrices for `size` z-stacked ranks of an nx*ny*nz block."""

    def __init__(self, nx: int, ny: int, nz: int, size: int = 1, stencil: int = 27, variant: str | None = None):
        if variant is None:
            variant = "mpi" if size > 1 else "serial"
        self.variant = variant
        self.lib = _lib(variant)
        self.size = size
        self.dims = (nx, ny, nz)
        self.h = self.lib.ref_create(nx, ny, nz, size, 1 if stencil == 7 else 0)
        if not self.h:
            raise RuntimeError(f"ref_create failed (variant {variant}, size {size})")

    @classmethod
    def from_file(cls, path: str, size: int = 1, variant: str | None = None):
        """The reference's read_HPC_row (read_HPC_row.cpp:217-373) on `size` ranks (+ make_local_matrix when size > 1)."""
        self = cls.__new__(cls)
        self.variant = variant or ("mpi" if size > 1 else "serial")
        self.lib = _lib(self.variant)
        self.size = size
        self.dims = None
        self.h = self.lib.ref_create_from_file(str(path).encode(), size)
        if not self.h:
            raise RuntimeError("ref_create_from_file failed")
        return self

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def threads(self) -> int:
        return self.lib.ref_threads()

    def scalar(self, rank: int, name: str) -> int:
        v = self.lib.ref_scalar(self.h, rank, name.encode())
        return int(v)

    def array(self, rank: int, name: str) -> np.ndarray:
        n = self.lib.ref_array(self.h, rank, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=_DTYPES[name])
        self.lib.ref_array(self.h, rank, name.encode(), out.ctypes.data_as(C.c_void_p), n)
        return out

    def spmv(self, xs, exchange: bool = True, reps: int = 1):
        """xs[rank]: local_ncol doubles (halo tail overwritten when exchange). Returns ys."""
        ys = [np.empty(self.scalar(r, "local_nrow")) for r in range(self.size)]
        self.lib.ref_spmv(self.h, _ppd(xs), _ppd(ys), int(exchange), reps)
        return ys

    def ddot(self, xs, ys):
        res = np.zeros(self.size)
        self.lib.ref_ddot(self.h, _ppd(xs), _ppd(ys), _pd(res))
        return res

    def solve(self, max_iter: int = 150, tol: float = 0.0, hist: bool = True, want_x: bool = True):
        h = np.full(max_iter, np.nan)
        niters = C.c_int(0)
        normr = C.c_double(0.0)
        times = np.zeros(7)
        xs = [np.empty(self.scalar(r, "local_nrow")) for r in range(self.size)] if want_x else None
        self.lib.ref_solve(self.h, max_iter, tol, int(hist), _pd(h), C.byref(niters), C.byref(normr), _pd(times),
                           _ppd(xs) if want_x else None)
        return {"hist": h, "niters": niters.value, "normr": normr.value, "times": times, "x": xs}

    def compute_residual(self, xs):
        res = np.zeros(self.size)
        self.lib.ref_compute_residual(self.h, _ppd(xs), _pd(res))
        return res


def waxpby(alpha: float, x: np.ndarray, beta: float, y: np.ndarray, w: np.ndarray | None = None,
           variant: str = "serial") -> np.ndarray:
    if w is None:
        w = np.empty_like(x)
    _lib(variant).ref_waxpby(len(x), alpha, _pd(x), beta, _pd(y), _pd(w))
    return w


def ddot_raw(x: np.ndarray, y: np.ndarray, variant: str = "serial") -> float:
    r = C.c_double(0.0)
    rc = _lib(variant).ref_ddot_raw(len(x), _pd(x), _pd(y), C.cast(C.byref(r), _PD))
    assert rc == 0
    return r.value


def yaml_report(nx, ny, nz, niters, normr, times, total_nrow, total_nnz, ranks=0, omp_threads=0, t4stats=None,
                variant: str = "serial", cwd: str | None = None) -> str:
    times = np.ascontiguousarray(times, dtype=np.float64)
    t4 = np.ascontiguousarray(t4stats if t4stats is not None else [0.0, 0.0, 0.0], dtype=np.float64)
    buf = C.create_string_buffer(1 << 16)
    old = os.getcwd()
    if cwd:
        os.chdir(cwd)
    try:
        n = _lib(variant).ref_yaml_report(nx, ny, nz, niters, normr, _pd(times), float(total_nrow), float(total_nnz),
                                          ranks, omp_threads, _pd(t4), buf, len(buf))
    finally:
        os.chdir(old)
    assert n >= 0
    return buf.value.decode()
