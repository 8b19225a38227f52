"""-m "not gpu": bench.py's reference arm (the reference's own CPU build, oracle/_ref) prints the contract's JSON line,
alone and under torchrun (rank 0 prints, the other ranks exit 0 without work)."""
import json
import socket
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def check_line(out, n_gpus):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "cg_gflops" and d["unit"] == "GFLOP/s" and d["n_gpus"] == n_gpus
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    return d


def test_reference_arm_single(H):
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--nx", "32", "--ny", "32", "--nz", "32",
           "--steps", "2", "--warmup", "1", "--cpu-iters", "10"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    d = check_line(res.stdout, 1)
    assert d["steps"] == 2 and d["warmup"] == 1


def test_reference_arm_under_torchrun(H):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--nx", "32",
           "--ny", "32", "--nz", "16", "--steps", "1", "--warmup", "1", "--cpu-iters", "5"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    d = check_line(res.stdout, 2)
    import os
    # torchrun's OMP_NUM_THREADS=1 default must not throttle the reference: all host threads, as at N = 1
    if d["cpu_baseline"]["variant"] == "omp":
        assert d["cpu_baseline"]["cores"] == os.cpu_count()


def test_gpu_arm_refuses_without_gpu(H):
    import torch
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--nx", "8", "--ny", "8", "--nz", "8"], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode != 0 and "no CPU fallback" in (res.stdout + res.stderr)
