"""-m gpu: the reference-compatible command-line driver (hpccg-sycl_b200/apps/test_HPCCG.cpp):
`test_HPCCG nx ny nz` prints the reference's residual lines and YAML keys (main.cpp:230-304, out.txt) and writes
./hpccg-1.0_<timestamp>.yaml (YAML_Doc.cpp:49-70)."""
import os
import re
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "hpccg-sycl_b200" / "lib" / "test_HPCCG"


def run(args, cwd):
    return subprocess.run([str(EXE), *args], cwd=cwd, capture_output=True, text=True, timeout=300)


def test_cli_matches_out_txt(H, cuda, tmp_path):
    """out.txt of the reference: 10x10x10 serial, 149 iterations, 'Initial Residual = 258.24',
    'Iteration = 15   Residual = 2.15402e-06', FLOP counts 9.536e+06 / 596000 / 894000 / 8.046e+06."""
    res = run(["10", "10", "10", "--check"], tmp_path)
    assert res.returncode == 0, res.stdout + res.stderr
    out = res.stdout
    assert "Initial Residual = 258.24\n" in out
    assert "Iteration = 15   Residual = 2.15402e-06\n" in out
    assert "Number of iterations: 149\n" in out
    for line in ("Mini-Application Name: hpccg", "Mini-Application Version: 1.0", "  nx: 10", "  Total   : 9.536e+06",
                 "  DDOT    : 596000", "  WAXPBY  : 894000", "  SPARSEMV: 8.046e+06", "Time Summary: ", "MFLOPS Summary: "):
        assert line + "\n" in out, line
    m = re.search(r"Difference between computed and exact: (\S+)", out)
    assert m and float(m.group(1)) <= 1e-12
    assert [f for f in os.listdir(tmp_path) if re.fullmatch(r"hpccg-1\.0_\d{4}_\d\d_\d\d__\d\d_\d\d_\d\d\.yaml", f)]


def test_cli_7pt_device_only_and_usage(H, cuda, tmp_path):
    res = run(["32", "32", "32", "--stencil", "7", "--device-only", "--iters", "40", "--check"], tmp_path)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Number of iterations: 39\n" in res.stdout and "Stencil points: 7" in res.stdout
    assert run([], tmp_path).returncode == 1
    assert "Usage:" in run(["4", "4"], tmp_path).stderr


def test_cli_ranks_mode(H, cuda, tmp_path):
    """`test_HPCCG nx ny nz --ranks N` = the reference's `mpirun -np N` run: one process per GPU, rank 0 prints the report
    with the MPI-only blocks (main.cpp:284-298).  On a box with fewer GPUs than ranks it must refuse (two ranks on one GPU
    would wait for each other inside kernels)."""
    ngpu = cuda.cuda.device_count()
    res = run(["16", "16", "8", "--ranks", str(ngpu + 1)], tmp_path)
    assert res.returncode == 3 and "needs" in res.stderr
    if ngpu < 2:
        pytest.skip("needs 2 GPUs for the positive case")
    res = run(["32", "32", "16", "--ranks", "2", "--check"], tmp_path)
    assert res.returncode == 0, res.stdout + res.stderr
    out = res.stdout
    assert "Number of MPI ranks: 2\n" in out and "Number of iterations: 149\n" in out
    assert "DDOT Timing Variations: \n" in out and "SPARSEMV OVERHEADS: \n" in out
    assert out.count("Initial Residual = ") == 2  # rank 0 only, once per solve (the driver solves twice)
    m = re.search(r"Difference between computed and exact: (\S+)", out)
    assert m and float(m.group(1)) <= 1e-12


def test_cli_mode2_matrix_file(H, refwrap, cuda, tmp_path):
    """`test_HPCCG HPC_data_file` (main.cpp:160-167): the reference's 10x10x10 matrix written in read_HPC_row's format gives
    the same residual lines as the generated one (out.txt:1-2)."""
    from test_read_hpc_row import stencil_file
    path, _ = stencil_file(refwrap, tmp_path, dims=(10, 10, 10))
    res = run([str(path), "--check"], tmp_path)
    assert res.returncode == 0, res.stdout + res.stderr
    assert f"Reading matrix info from {path}..." in res.stdout
    assert "Initial Residual = 258.24\n" in res.stdout and "Iteration = 15   Residual = 2.15402e-06\n" in res.stdout
    assert "Number of iterations: 149\n" in res.stdout
    assert run([str(tmp_path / "nope.dat")], tmp_path).returncode == 1


def test_reference_main_cpp_runs_on_the_b200_library(H, refwrap, cuda, tmp_path):
    """Drop-in proof: oracle/_ref/test_HPCCG_refmain is the reference's OWN, unmodified main.cpp (read where it lies under
    /root/reference by oracle/build.sh) compiled against this repo's headers (the reference's header names) and linked with
    libhpccg_b200.so instead of the reference's kernels.  main.cpp hard-codes max_iter = 500 (main.cpp:187)."""
    exe = ROOT / "oracle" / "_ref" / "test_HPCCG_refmain"
    if not exe.exists():
        pytest.skip("oracle/_ref/test_HPCCG_refmain not built (needs /root/reference at build time)")
    res = subprocess.run([str(exe), "64", "64", "64"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    out = res.stdout
    with refwrap.RefWorld(64, 64, 64, variant="serial" if refwrap.available("serial") else "oracle") as R:
        ref = R.solve(500, hist=False)
        href = R.solve(500, hist=True)["hist"]
    assert "Error in call to CG" not in res.stderr
    assert "Initial Residual = 1654.81\n" in out                       # 1654.8087502790163
    m = re.search(r"Iteration = 50   Residual = (\S+)", out)            # print_freq = 50 at max_iter = 500 (HPCCG.cpp:342-343)
    assert m and abs(float(m.group(1)) - href[50]) <= 1e-5 * href[50]  # 6 printed digits
    m = re.search(r"Number of iterations: (\d+)", out)
    assert m and int(m.group(1)) == ref["niters"] == 499
    assert "  nx: 64\n" in out and "MFLOPS Summary: \n" in out and "  SPARSEMV: " in out
