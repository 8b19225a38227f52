"""-m "not gpu": the C-ABI library loads, exports every symbol include/hpccg_b200.h declares, the ctypes table
covers exactly that set, and the no-GPU error behaviour is loud (non-zero return + message, never a CPU result)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "hpccg_b200.h"


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    names = re.findall(r"\b(hpccg_[a-zA-Z0-9_]+)\s*\(", text)
    return sorted(set(n for n in names if n not in ("hpccg_allgather_fn",)))


def test_header_declares_functions():
    names = header_functions()
    assert len(names) >= 55
    for must in ("hpccg_dev_spmv", "hpccg_dev_dot", "hpccg_dev_waxpby", "hpccg_dev_spmv_dot", "hpccg_dev_update_xr_dot",
                 "hpccg_dev_p_update", "hpccg_dev_halo_pack", "hpccg_dev_cg_solve", "hpccg_api_HPCCG",
                 "hpccg_api_HPC_sparsemv", "hpccg_api_ddot", "hpccg_api_waxpby", "hpccg_api_generate_matrix",
                 "hpccg_api_make_local_matrix", "hpccg_api_exchange_externals", "hpccg_api_yaml_report"):
        assert must in names


def test_library_exports_every_declared_symbol(H):
    from hpccg_sycl_b200 import _capi
    out = subprocess.run(["nm", "-D", "--defined-only", str(_capi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if " T " in line)
    declared = header_functions()
    missing = [n for n in declared if n not in exported]
    assert not missing, missing
    # the ctypes table binds exactly the header's functions
    assert sorted(_capi.SIGNATURES) == declared
    # no torch / C++ types cross the boundary: every exported hpccg_* symbol is unmangled
    assert not [s for s in exported if "hpccg_" in s and s.startswith("_Z") and "hpccg_dev_matrix" not in s and "hpccg5" not in s
                and "hpccg" == s[:5]]


def test_reference_named_cxx_api_is_exported(H):
    """The reference's own C++ signatures are present with C++ linkage."""
    from hpccg_sycl_b200 import _capi
    out = subprocess.run(["nm", "-DC", "--defined-only", str(_capi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    for sig in ("HPCCG(HPC_Sparse_Matrix_STRUCT*, double*, double*, int, double, int&, double&, double*)",
                "HPC_sparsemv(HPC_Sparse_Matrix_STRUCT*, double const*, double*)",
                "ddot(int, double const*, double const*, double*, double&)",
                "waxpby(int, double, double const*, double, double const*, double*)",
                "generate_matrix(int, int, int, HPC_Sparse_Matrix_STRUCT**, double**, double**, double**)",
                "make_local_matrix(HPC_Sparse_Matrix_STRUCT*)",
                "exchange_externals(HPC_Sparse_Matrix_STRUCT*, double const*)",
                "destroyMatrix(HPC_Sparse_Matrix_STRUCT*&)",
                "compute_residual(int, double const*, double const*, double*)",
                "mytimer()", "YAML_Doc::generateYAML", "YAML_Element::get"):
        assert sig in out, sig


def test_sm100a_cubin_embedded(H):
    from hpccg_sycl_b200 import _capi
    out = subprocess.run(["cuobjdump", "-lelf", str(_capi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_version_and_context(H):
    from hpccg_sycl_b200._capi import lib
    assert lib.hpccg_version() == 100
    H.set_rank(2, 5)
    assert H.get_rank() == (2, 5)
    H.set_rank(0, 1)
    with pytest.raises(H.HpccgError):
        H.set_rank(3, 3)
    with pytest.raises(H.HpccgError):
        H.set_options(9, True)


def test_no_cpu_fallback_without_gpu(H):
    """Without a usable GPU every compute entry point FAILS with a CUDA error; nothing is computed on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path cannot be exercised")
    H.set_rank(0, 1)
    H.set_options(27, True)
    A = H.generate_matrix(4, 4, 4)   # host set-up is allowed: it is bit-exact index/byte work of the boundary
    n = A.local_nrow
    x, y = np.ones(n), np.full(n, np.nan)
    with pytest.raises(H.HpccgError, match="CUDA error"):
        H.HPC_sparsemv(A, x, y)
    assert np.isnan(y).all()
    with pytest.raises(H.HpccgError, match="CUDA error"):
        H.ddot(n, x, x)
    with pytest.raises(H.HpccgError, match="CUDA error"):
        H.waxpby(n, 1.0, x, 2.0, x, y)
    assert np.isnan(y).all()
    with pytest.raises(H.HpccgError, match="CUDA error"):
        H.HPCCG(A, A.b, A.x.copy(), 10, 0.0)
    A.destroy()


def test_missing_library_is_an_import_error(tmp_path):
    """The Python layer has no fallback: without libhpccg_b200.so the import itself fails."""
    code = (
        "import sys, importlib.util, pathlib\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "import hpccg_pkg\n"
        "from pathlib import Path\n"
        "import shutil\n"
        f"dst = Path({str(tmp_path)!r}) / 'pkg'\n"
        "shutil.copytree(hpccg_pkg.PKG_DIR, dst, ignore=shutil.ignore_patterns('lib', '__pycache__'))\n"
        "hpccg_pkg.PKG_DIR = dst\n"
        "try:\n"
        "    hpccg_pkg.load()\n"
        "except ImportError as e:\n"
        "    print('IMPORT-ERROR', 'no Python or CPU fallback' in str(e))\n"
    )
    out = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert "IMPORT-ERROR True" in out.stdout, out.stdout + out.stderr
