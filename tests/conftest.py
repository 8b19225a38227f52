import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_once():
    # The library and the oracle are built in-tree by __graft_entry__.build(); do it here too so that a
    # bare `pytest` on a fresh checkout works (seconds when up to date).
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def H():
    _build_once()
    import hpccg_pkg
    return hpccg_pkg.load()


@pytest.fixture(scope="session")
def refwrap(H):
    import refwrap as r
    return r


def ref_variant(size=1):
    """The strongest available checker: the real reference where oracle/_ref has it, else the C restatement."""
    import refwrap as r
    want = "mpi" if size > 1 else "serial"
    return want if r.available(want) else "oracle"


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch
