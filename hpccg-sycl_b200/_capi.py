"""ctypes declarations for every symbol of include/hpccg_b200.h.

There is no Python fallback: if lib/libhpccg_b200.so is missing the import fails with the build
command, and every wrapper raises on a non-zero return code with the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libhpccg_b200.so"

PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int)
PLL = C.POINTER(C.c_longlong)
VP = C.c_void_p
PVP = C.POINTER(C.c_void_p)

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p)

# name -> (restype, argtypes); the list is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "hpccg_last_error": (C.c_char_p, []),
    "hpccg_version": (C.c_int, []),
    "hpccg_device_count": (C.c_int, [PI]),
    "hpccg_set_device": (C.c_int, [C.c_int]),
    "hpccg_device_synchronize": (C.c_int, []),
    "hpccg_dev_malloc": (C.c_int, [PVP, C.c_longlong]),
    "hpccg_dev_free": (C.c_int, [VP]),
    "hpccg_host_malloc_pinned": (C.c_int, [PVP, C.c_longlong]),
    "hpccg_host_free_pinned": (C.c_int, [VP]),
    "hpccg_memcpy_h2d": (C.c_int, [VP, VP, C.c_longlong, VP]),
    "hpccg_memcpy_d2h": (C.c_int, [VP, VP, C.c_longlong, VP]),
    "hpccg_stream_synchronize": (C.c_int, [VP]),
    "hpccg_ctx_set": (C.c_int, [C.c_int, C.c_int]),
    "hpccg_ctx_get": (C.c_int, [PI, PI]),
    "hpccg_ctx_set_allgather": (C.c_int, [ALLGATHER_FN, VP]),
    "hpccg_local_world_create": (C.c_int, [C.c_int, PVP]),
    "hpccg_local_world_bind": (C.c_int, [VP, C.c_int]),
    "hpccg_local_world_destroy": (C.c_int, [VP]),
    "hpccg_nccl_available": (C.c_int, []),
    "hpccg_nccl_unique_id": (C.c_int, [VP]),
    "hpccg_nccl_init": (C.c_int, [VP, C.c_int, C.c_int]),
    "hpccg_nccl_finalize": (C.c_int, []),
    "hpccg_dev_matrix_create": (C.c_int, [C.c_int, C.c_int, VP, VP, VP, PVP]),
    "hpccg_dev_matrix_generate": (C.c_int, [C.c_int] * 6 + [VP, VP, C.c_int, PVP]),
    "hpccg_dev_matrix_set_halo": (C.c_int, [VP, C.c_int, VP, VP, VP, VP, C.c_int]),
    "hpccg_dev_matrix_destroy": (C.c_int, [VP]),
    "hpccg_dev_matrix_info": (C.c_int, [VP, PI, PI, PI, PLL]),
    "hpccg_dev_matrix_download": (C.c_int, [VP, VP, VP]),
    "hpccg_dev_matrix_bytes": (C.c_int, [VP, PLL]),
    "hpccg_dev_matrix_compress": (C.c_int, [VP]),
    "hpccg_dev_matrix_format": (C.c_int, [VP, PI, PI]),
    "hpccg_dev_matrix_comm": (C.c_int, [VP, PI, PI]),
    "hpccg_dev_spmv": (C.c_int, [VP, VP, VP, VP]),
    "hpccg_dev_dot": (C.c_int, [C.c_int, VP, VP, VP, VP]),
    "hpccg_dev_waxpby": (C.c_int, [C.c_int, C.c_double, VP, C.c_double, VP, VP, VP]),
    "hpccg_dev_spmv_dot": (C.c_int, [VP, VP, VP, VP, VP]),
    "hpccg_dev_update_xr_dot": (C.c_int, [C.c_int, VP, VP, VP, VP, VP, VP, VP]),
    "hpccg_dev_p_update": (C.c_int, [C.c_int, VP, VP, VP, VP]),
    "hpccg_dev_halo_pack": (C.c_int, [VP, VP, VP, VP]),
    "hpccg_dev_max_abs_diff": (C.c_int, [C.c_int, VP, VP, VP, VP]),
    "hpccg_dev_cg_solve": (C.c_int, [VP, VP, VP, C.c_int, C.c_double, PI, PD, VP, VP, PD, C.c_int, VP]),
    "hpccg_dev_cg_solve_group": (C.c_int, [C.c_int, PVP, PVP, PVP, C.c_int, C.c_double, PI, PD, VP, PD, C.c_int, VP]),
    "hpccg_launch_count": (C.c_longlong, []),
    "hpccg_api_set_options": (C.c_int, [C.c_int, C.c_int]),
    "hpccg_api_set_print": (C.c_int, [C.c_int]),
    "hpccg_api_set_matrix_format": (C.c_int, [C.c_int]),
    "hpccg_api_generate_matrix": (C.c_int, [C.c_int, C.c_int, C.c_int, PVP, C.POINTER(PD), C.POINTER(PD), C.POINTER(PD)]),
    "hpccg_api_read_HPC_row": (C.c_int, [C.c_char_p, PVP, C.POINTER(PD), C.POINTER(PD), C.POINTER(PD)]),
    "hpccg_api_make_local_matrix": (C.c_int, [VP]),
    "hpccg_api_HPCCG": (C.c_int, [VP, VP, VP, C.c_int, C.c_double, PI, PD, PD]),
    "hpccg_api_HPC_sparsemv": (C.c_int, [VP, VP, VP]),
    "hpccg_api_ddot": (C.c_int, [C.c_int, VP, VP, PD, PD]),
    "hpccg_api_waxpby": (C.c_int, [C.c_int, C.c_double, VP, C.c_double, VP, VP]),
    "hpccg_api_exchange_externals": (C.c_int, [VP, VP]),
    "hpccg_api_compute_residual": (C.c_int, [C.c_int, VP, VP, PD]),
    "hpccg_api_destroyMatrix": (C.c_int, [VP]),
    "hpccg_api_free_vectors": (C.c_int, [PD, PD, PD]),
    "hpccg_api_matrix_scalar": (C.c_longlong, [VP, C.c_char_p]),
    "hpccg_api_matrix_array": (C.c_longlong, [VP, C.c_char_p, VP, C.c_longlong]),
    "hpccg_api_matrix_device": (C.c_int, [VP, PVP]),
    "hpccg_api_last_history": (C.c_int, [PD, C.c_int]),
    "hpccg_api_yaml_report": (C.c_int, [C.c_int] * 4 + [C.c_double, PD, C.c_double, C.c_double, C.c_int, C.c_int, PD,
                                                        C.c_char_p, C.c_int]),
}


class HpccgError(RuntimeError):
    pass


def load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no Python or CPU fallback for the HPCCG kernels.")
    # RTLD_LOCAL: the library exports the reference's C++ names (HPCCG, ddot, ...); keep them out of the
    # global scope so that the oracle (the real reference, same names) can live in the same test process.
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library drift
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise HpccgError(f"{what or 'hpccg'} failed with code {rc}: {lib.hpccg_last_error().decode(errors='replace')}")
