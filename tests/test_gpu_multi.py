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
    for r in range(world):
        d = json.loads((tmp_path / f"rank{r}.json").read_text())
        assert d["size"] == world
        for c in d["cases"]:
            assert c["worst_rel"] <= 1e-8 and c["ddot_rel"] <= 1e-12
            assert c["residual"] <= 1e-12 or c["residual"] != c["residual"], c  # NaN only where the reference has it
