"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): the real multi-GPU path -- one process per GPU under
torchrun, NCCL halo exchange overlapped with the interior SpMV, NCCL scalar gathers -- against the oracle's
N-rank world.  (The same numerics are covered on ONE GPU by test_gpu_solve.py's in-process group solve.)"""
import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_solve_matches_reference(cuda, tmp_path, world):
    if cuda.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(HERE / "_nccl_worker.py"), str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    summary = []
    for r in range(world):
        d = json.loads((tmp_path / f"rank{r}.json").read_text())
        assert d["size"] == world
        for c in d["cases"]:
            assert c["worst_rel"] <= 1e-8 and c["ddot_rel"] <= 1e-12
            assert c["residual"] <= 1e-12 or c["residual"] != c["residual"], c  # NaN only where the reference has it
        if r == 0:
            summary = d["cases"]
    # keep a record of what ran (gpurun_out/ travels back from the GPU box; copied to profiles/ by hand)
    out = HERE.parent / "gpurun_out"
    if out.is_dir():
        (out / f"multi_gpu_parity_world{world}.json").write_text(json.dumps(
            {"world": world, "cases": len(summary), "worst_rel": max(c["worst_rel"] for c in summary),
             "worst_ddot_rel": max(c["ddot_rel"] for c in summary), "per_case": summary}, indent=1))
