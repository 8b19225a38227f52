"""-m "not gpu": pins the oracle.

The C restatement (oracle/hpccg_oracle.c) is checked bit for bit against
  (1) the committed golden fixtures, which tests/golden/make_golden.py generated from the REAL reference
      (unmodified /root/reference sources compiled by oracle/build.sh), and
  (2) the real reference itself wherever oracle/_ref/libhpccg_ref_*.so exists (always in the build
      container; on the GPU box the prebuilt libraries travel with the snapshot).
It also pins the few numbers the reference tree itself holds for this path (out.txt).
"""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())
ARRAYS = ["nnz_in_row", "list_of_inds", "list_of_vals", "ind_offsets", "val_offsets", "diag_offsets", "x", "b", "xexact",
          "external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seeded(n, seed):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


def cfg_id(rec):
    return "x".join(map(str, rec["dims"])) + f"-r{rec['ranks']}-s{rec['stencil']}"


@pytest.mark.parametrize("rec", GOLDEN["configs"], ids=cfg_id)
def test_restatement_matches_golden(refwrap, rec):
    nx, ny, nz = rec["dims"]
    size = rec["ranks"]
    with refwrap.RefWorld(nx, ny, nz, size=size, stencil=rec["stencil"], variant="oracle") as R:
        for r in range(size):
            g = rec["rank"][r]
            for s, v in g["scalars"].items():
                assert R.scalar(r, s) == v, (r, s)
            for a in ARRAYS:
                assert sha(R.array(r, a)) == g["sha256"][a], (r, a)
            for a, v in g["small"].items():
                assert R.array(r, a).tolist() == v
        nrow = [R.scalar(r, "local_nrow") for r in range(size)]
        ncol = [R.scalar(r, "local_ncol") for r in range(size)]
        xs = [np.concatenate([seeded(nrow[r], 12345 + r), np.zeros(ncol[r] - nrow[r])]) for r in range(size)]
        ys = R.spmv(xs, exchange=True)
        assert [sha(y) for y in ys] == rec["spmv_sha256"]
        v2 = [seeded(nrow[r], 54321 + r) for r in range(size)]
        xloc = [x[:nrow[r]].copy() for r, x in enumerate(xs)]
        assert float(R.ddot(xloc, v2)[0]).hex() == rec["ddot_xy"]
        assert float(R.ddot(xloc, xloc)[0]).hex() == rec["ddot_xx"]
        s = R.solve(rec["max_iter"], 0.0, hist=True)
        assert s["niters"] == rec["niters"]
        assert float(s["normr"]).hex() == rec["normr"]
        assert [float(v).hex() for v in s["hist"]] == rec["hist"]  # every iteration, every bit


def test_restatement_waxpby_matches_golden(refwrap):
    w = GOLDEN["waxpby"]
    x, y = seeded(w["n"], 12345), seeded(w["n"], 54321)
    for name, c in w["cases"].items():
        assert sha(refwrap.waxpby(c["alpha"], x, c["beta"], y, variant="oracle")) == c["sha256"], name


LIVE = [(20, 30, 10, 1, 27), (20, 30, 10, 1, 7), (9, 7, 3, 1, 27), (6, 5, 2, 3, 27), (6, 5, 1, 4, 27), (8, 8, 4, 2, 7),
        (24, 24, 6, 5, 27)]


@pytest.mark.parametrize("nx,ny,nz,size,stencil", LIVE)
def test_restatement_matches_live_reference(refwrap, nx, ny, nz, size, stencil):
    """Shapes that are NOT in the fixtures, against the real reference run here."""
    variant = "mpi" if size > 1 else "serial"
    if not refwrap.available(variant):
        pytest.skip("real reference not built (no /root/reference and no prebuilt oracle/_ref)")
    with refwrap.RefWorld(nx, ny, nz, size=size, stencil=stencil, variant=variant) as R, \
            refwrap.RefWorld(nx, ny, nz, size=size, stencil=stencil, variant="oracle") as O:
        for r in range(size):
            for s in refwrap.SCALARS:
                assert O.scalar(r, s) == R.scalar(r, s), (r, s)
            for a in ARRAYS:
                assert np.array_equal(O.array(r, a), R.array(r, a)), (r, a)
        ncol = [R.scalar(r, "local_ncol") for r in range(size)]
        xs = [seeded(ncol[r], 99 + r) for r in range(size)]
        yo = O.spmv([x.copy() for x in xs])
        yr = R.spmv([x.copy() for x in xs])
        for a, b in zip(yo, yr):
            assert np.array_equal(a, b)
        so, sr = O.solve(60), R.solve(60)
        assert so["niters"] == sr["niters"]
        assert np.array_equal(so["hist"], sr["hist"], equal_nan=True)
        for a, b in zip(so["x"], sr["x"]):
            assert np.array_equal(a, b)


def test_reference_out_txt_values(refwrap):
    """The only numbers the reference tree holds for this path: out.txt:1-2,21 (10x10x10 serial:
    'Initial Residual = 258.24', 'Iteration = 15   Residual = 2.15402e-06', 149 iterations) and its FLOP
    counts (out.txt:29-32 = main.cpp:217-227 with nrow = 1000, nnz = 27000, 149 iterations)."""
    with refwrap.RefWorld(10, 10, 10, variant="oracle") as R:
        s = R.solve(150)
    assert s["niters"] == 149
    assert f"{s['hist'][0]:.6g}" == "258.24"
    assert f"{s['hist'][15]:.6g}" == "2.15402e-06"
    it, nrow, nnz = 149.0, 1000.0, 27000.0
    assert (it * 4 * nrow, it * 6 * nrow, it * 2 * nnz) == (596000.0, 894000.0, 8.046e6)
    assert it * (4 * nrow + 6 * nrow + 2 * nnz) == 9.536e6


def test_closed_form_structure(refwrap):
    """real nnz = (3nx-2)(3ny-2)(3nz_g-2) for 27-pt, 7n - 2(nxny+nynz+nxnz) for 7-pt;
    b = 27 - (nnz_row - 1) = A*1; xexact = 1."""
    for (nx, ny, nz, size) in ((20, 30, 10, 1), (6, 5, 2, 3)):
        with refwrap.RefWorld(nx, ny, nz, size=size, variant="oracle") as R:
            tot = sum(R.scalar(r, "nnz_sum") for r in range(size))
            assert tot == (3 * nx - 2) * (3 * ny - 2) * (3 * nz * size - 2)
            for r in range(size):
                assert np.array_equal(R.array(r, "b"), 27.0 - (R.array(r, "nnz_in_row") - 1.0))
    with refwrap.RefWorld(20, 30, 10, stencil=7, variant="oracle") as R:
        assert R.scalar(0, "nnz_sum") == 7 * 6000 - 2 * (600 + 300 + 200) == 39800


def test_known_answer_x_converges_to_one(refwrap):
    with refwrap.RefWorld(20, 30, 10, variant="oracle") as R:
        s = R.solve(150)
        assert R.compute_residual(s["x"])[0] <= 1e-12
