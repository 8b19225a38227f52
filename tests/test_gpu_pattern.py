"""-m gpu: the opt-in pattern-coded matrix format (hpccg_dev_matrix_compress).  It is a lossless
re-encoding -- one 16-bit id per row into a table of distinct row patterns (sequences of (value, column - row) pairs)
-- so every result must be BIT-IDENTICAL to the default SELL format and to the reference."""
import ctypes as C

import numpy as np
import pytest

from conftest import ref_variant
from test_gpu_solve import check_history, check_solution, _build_ranks

pytestmark = pytest.mark.gpu


def seeded(n, seed):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


@pytest.fixture()
def dict_format(H):
    H.set_matrix_format("pattern")
    yield
    H.set_matrix_format("sell")


@pytest.mark.parametrize("dims,stencil", [((20, 30, 10), 27), ((20, 30, 10), 7), ((33, 17, 5), 27), ((64, 64, 8), 27),
                                           ((3, 3, 3), 27), ((128, 4, 4), 7)])
def test_sparsemv_bit_exact_and_mirror_identical(H, refwrap, cuda, dims, stencil):
    H.set_rank(0, 1)
    H.set_options(stencil, True)
    A = H.generate_matrix(*dims)
    m = A.device()
    slots = m.info()["slots"]
    v0, c0 = m.download()
    bytes0 = m.bytes()
    f = m.compress()
    if slots in (27, 7):
        # a 27-pt block has at most 3*3*3 boundary types, a 7-pt block the same (each of x, y, z: first / inner / last)
        assert f["format"] == 1 and 1 <= f["patterns"] <= 27
        assert m.bytes() < bytes0 / 6
    else:
        assert f["format"] == 0  # thin blocks (fewer slots) have no pattern kernel: left as they are
    v1, c1 = m.download()
    assert np.array_equal(c0, c1) and np.array_equal(v0, v1)
    n = A.local_nrow
    x = seeded(n, 12345)
    y = np.full(n, np.nan)
    H.HPC_sparsemv(A, x, y)
    with refwrap.RefWorld(*dims, stencil=stencil, variant=ref_variant()) as R:
        yr = R.spmv([x.copy()])[0]
    assert np.array_equal(y, yr)
    A.destroy()


@pytest.mark.parametrize("dims,stencil", [((20, 30, 10), 27), ((64, 64, 64), 27), ((32, 32, 32), 7)])
def test_hpccg_dict_format_matches_reference_and_sell_bitwise(H, refwrap, cuda, dict_format, dims, stencil):
    H.set_rank(0, 1)
    H.set_options(stencil, False)  # device-generated, then compressed at mirror creation
    A = H.generate_matrix(*dims)
    assert A.device().format()["format"] == 1
    x = A.x.copy()
    niters, normr, _, hist = H.HPCCG(A, A.b, x, 150, 0.0)
    with refwrap.RefWorld(*dims, stencil=stencil, variant=ref_variant()) as R:
        ref = R.solve(150)
    check_history(hist, ref["hist"], niters, ref["niters"])
    check_solution(x, ref["x"][0])
    A.destroy()
    # the same solve in the default format: identical bits, iteration by iteration (same kernels' arithmetic, same
    # reduction tree when the grids agree; otherwise within the reduction-order bar)
    H.set_matrix_format("sell")
    B = H.generate_matrix(*dims)
    xb = B.x.copy()
    nb, normb, _, histb = H.HPCCG(B, B.b, xb, 150, 0.0)
    check_history(hist, histb, niters, nb)
    B.destroy()
    H.set_options(stencil, True)


@pytest.mark.parametrize("dims,size,stencil", [((16, 16, 8), 2, 27), ((12, 10, 2), 3, 27), ((32, 32, 16), 4, 27), ((8, 8, 4), 2, 7)])
def test_multi_rank_dict_format_with_halo_columns(H, refwrap, cuda, dict_format, dims, size, stencil):
    """Rows that reference halo columns (>= local_nrow) have their own patterns; the two plane rows whose halo numbering
    is interleaved (first-encounter order) get one pattern per row."""
    torch = cuda
    mats = _build_ranks(H, dims, size, stencil, True)
    ms = [A.device() for A in mats]
    fmts = [m.format() for m in ms]
    if ms[0].info()["slots"] in (27, 7):
        assert all(f["format"] == 1 for f in fmts)
    bs = [torch.from_numpy(A.b.copy()).cuda() for A in mats]
    xs = [torch.zeros(A.local_nrow, dtype=torch.float64, device="cuda") for A in mats]
    out = H.dev.cg_solve_group(ms, bs, xs, 150, 0.0)
    with refwrap.RefWorld(*dims, size=size, stencil=stencil, variant=ref_variant(size)) as R:
        ref = R.solve(150)
        # per-rank SpMV with the reference's halo values
        ncol = [R.scalar(r, "local_ncol") for r in range(size)]
        nrow = [R.scalar(r, "local_nrow") for r in range(size)]
        xv = [np.concatenate([seeded(nrow[r], 7 + r), np.zeros(ncol[r] - nrow[r])]) for r in range(size)]
        ys = R.spmv(xv, exchange=True)  # fills the halo tails of xv
    check_history(out["hist"], ref["hist"], out["niters"], ref["niters"])
    for r, (A, m) in enumerate(zip(mats, ms)):
        xd = torch.from_numpy(xv[r]).cuda()
        yd = torch.empty(nrow[r], dtype=torch.float64, device="cuda")
        H.dev.spmv(m, xd, yd)
        assert np.array_equal(yd.cpu().numpy(), ys[r]), r
    for A in mats:
        A.destroy()


def _create_from_rows(H, nnz, vals, cols, ncol):
    """hpccg_dev_matrix_create from numpy row arrays (what the reference's struct holds)."""
    from hpccg_sycl_b200._capi import lib
    n = len(nnz)
    starts = np.concatenate([[0], np.cumsum(nnz)[:-1]]).astype(np.int64)
    pv = (C.c_void_p * n)(*[vals.ctypes.data + 8 * int(s) for s in starts])
    pc = (C.c_void_p * n)(*[cols.ctypes.data + 4 * int(s) for s in starts])
    out = C.c_void_p()
    nn = np.ascontiguousarray(nnz, dtype=np.int32)
    rc = lib.hpccg_dev_matrix_create(n, ncol, nn.ctypes.data, C.cast(pv, C.c_void_p), C.cast(pc, C.c_void_p), C.byref(out))
    assert rc == 0, lib.hpccg_last_error()
    return H.DeviceMatrix(out.value, True), (pv, pc, nn)


def _spmv_rows(nnz, vals, cols, x):
    y = np.zeros(len(nnz))
    k = 0
    for i, c in enumerate(nnz):
        s = 0.0
        for j in range(c):
            s = s + vals[k + j] * x[cols[k + j]]
        y[i] = s
        k += c
    return y


def test_perturbed_matrix_gets_more_patterns(H, refwrap, cuda):
    """The reference's 20x30x10 matrix with 64 rows perturbed: 64 more patterns, result still exact."""
    torch = cuda
    with refwrap.RefWorld(20, 30, 10, variant=ref_variant()) as R:
        nnz, vals, cols = R.array(0, "nnz_in_row"), R.array(0, "list_of_vals").copy(), R.array(0, "list_of_inds")
    rng = np.random.default_rng(3)
    starts = np.concatenate([[0], np.cumsum(nnz)[:-1]])
    m0, keep0 = _create_from_rows(H, nnz, vals, cols, 6000)
    base = m0.compress()["patterns"]
    m0.destroy()
    for row in list(range(0, 16)) + list(range(700, 716)) + list(range(3000, 3016)) + list(range(5984, 6000)):
        vals[starts[row]:starts[row] + nnz[row]] = rng.uniform(-2, 2, nnz[row])
    m, keep = _create_from_rows(H, nnz, vals, cols, 6000)
    v0, c0 = m.download()
    f = m.compress()
    # 64 new patterns; the few original patterns whose only rows were perturbed (the two corner rows) disappear
    assert f["format"] == 1 and base + 56 <= f["patterns"] <= base + 64, (f, base)
    v1, c1 = m.download()
    assert np.array_equal(v0, v1) and np.array_equal(c0, c1)
    x = seeded(6000, 11)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty(6000, dtype=torch.float64, device="cuda")
    H.dev.spmv(m, xd, yd)
    assert np.array_equal(yd.cpu().numpy(), _spmv_rows(nnz, vals, cols, x))
    m.destroy()


def test_incompressible_matrix_is_left_alone(H, cuda):
    """70 000 random rows = 70 000 patterns > 65535 ids: compress() leaves the matrix in format 0 (not an error)."""
    torch = cuda
    n = 70000
    rng = np.random.default_rng(5)
    nnz = np.full(n, 7, dtype=np.int32)
    vals = rng.uniform(-1, 1, 7 * n)
    cols = rng.integers(0, n, 7 * n).astype(np.int32)
    m, keep = _create_from_rows(H, nnz, vals, cols, n)
    assert m.compress()["format"] == 0
    x = seeded(n, 1)
    yd = torch.empty(n, dtype=torch.float64, device="cuda")
    H.dev.spmv(m, torch.from_numpy(x).cuda(), yd)
    ref = (vals.reshape(n, 7) * x[cols.reshape(n, 7)])
    acc = np.zeros(n)
    for j in range(7):
        acc = acc + ref[:, j]
    assert np.array_equal(yd.cpu().numpy(), acc)
    m.destroy()
