"""Build recipe for libhpccg_b200.so (explicit nvcc, sm_100a only, in-tree output).

The library is the product: hand-written CUDA kernels + the C-ABI of include/hpccg_b200.h + the
reference-named C++ API.  It is built IN-TREE (hpccg-sycl_b200/lib/) so that it travels to the GPU box
with the repository snapshot; nvcc cross-compiles here without a GPU.
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libhpccg_b200.so"
SOURCES = ["device_api.cu", "context.cu", "host_api.cu", "yaml.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xlinker", "-Bsymbolic",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + list(CSRC.glob("*.cpp")) + \
        list((CSRC / "include").glob("*.hpp")) + [PKG.parent / "include" / "hpccg_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into lib/libhpccg_b200.so."""
    if not force and not _stale():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    # the image exports CXX=/opt/gcc/bin/g++ (no libgomp, odd specs); pin the system host compiler
    cmd = [_nvcc(), "-ccbin", "/usr/bin/g++", *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(LIB), *[str(CSRC / s) for s in SOURCES], "-ldl"]
    res = subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def build_driver() -> Path:
    """The reference-compatible command-line driver (apps/test_HPCCG.cpp) linked against the library."""
    out = PKG / "lib" / "test_HPCCG"
    src = PKG / "apps" / "test_HPCCG.cpp"
    if not src.exists():
        return out
    if out.exists() and out.stat().st_mtime > max(src.stat().st_mtime, LIB.stat().st_mtime):
        return out
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", f"-I{CSRC / 'include'}", f"-I{PKG.parent / 'include'}", str(src),
           "-o", str(out), f"-L{LIBDIR}", "-lhpccg_b200", f"-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("driver build failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_driver())
