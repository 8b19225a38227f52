"""-m "not gpu": the N>1 host path on CPU -- world_size 2 and 3, gloo backend, one process per rank."""
import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_setup_over_gloo(H, tmp_path, world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(HERE / "_dist_worker.py"), str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    for r in range(world):
        d = json.loads((tmp_path / f"rank{r}.json").read_text())
        assert d["size"] == world and d["rank"] == r
        assert all(c["ok"] for c in d["cases"]), d
        assert d["max_reduce"] == float(world)
