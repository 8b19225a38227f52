"""-m gpu: the device-resident CG loop (HPCCG.cpp:312-402) against the oracle.

Residual-history bar (north_star): <= 1e-8 relative, since only the reduction order differs.  SURVEY.md
section 4.1 shows the reference disagrees with ITSELF beyond that once normr has dropped ~10 decades (small and
7-pt problems), so the 1e-8 bar applies while normr_k >= 1e-10 * normr_0; after that the test requires that the residual
stays converged, and the known answer max|x-1| <= 1e-12 (or the reference's own NaN, see check_solution)."""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import ref_variant

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
REL_HIST = 1e-8


def reference_self_spread(dims, stencil=27, max_iter=150):
    """how far the reference is from ITSELF once the residual has dropped ten decades -- its serial build
    against its OpenMP build on the same problem (None when oracle/_ref has only one of them).  Returns (first iteration of
    the noise regime, largest relative difference inside the regular regime, largest decade distance beyond it)."""
    import refwrap
    if not (refwrap.available("serial") and refwrap.available("omp")):
        return None
    with refwrap.RefWorld(*dims, stencil=stencil, variant="serial") as R:
        a = R.solve(max_iter)["hist"]
    with refwrap.RefWorld(*dims, stencil=stencil, variant="omp") as R:
        b = R.solve(max_iter)["hist"]
    ran = ~np.isnan(a) & ~np.isnan(b) & (a > 0) & (b > 0)
    regular = ran & (a >= 1e-10 * a[0])
    noisy = ran & ~regular
    rel = float((np.abs(a[regular] - b[regular]) / a[regular]).max())
    decades = float(np.abs(np.log10(a[noisy] / b[noisy])).max()) if noisy.any() else 0.0
    first = int(np.argmax(noisy)) if noisy.any() else -1
    return first, rel, decades


def check_history(hist, ref_hist, niters, ref_niters):
    """Regular regime (normr_k >= 1e-10 normr_0): <= 1e-8 relative, every iteration.  Beyond it the recursion is
    rounding noise -- the reference's own serial / OpenMP / 3-thread builds are decades apart there (SURVEY.md 4.1;
    the first B200 run was 2.9 decades from the serial reference at 10^3, k > 40) -- so the test requires only
    that the residual stays converged (<= 1e-9 normr_0).  The iteration count must be equal unless the loop ended
    through exact underflow of r.r to 0 (HPCCG.cpp:358 with tolerance 0), which happens at a rounding-dependent
    iteration: then both runs must already be below 1e-140 at the shorter run's last iteration."""
    h0 = ref_hist[0]
    ran = ~np.isnan(ref_hist) & ~np.isnan(hist)
    regular = ran & (ref_hist >= 1e-10 * h0)
    assert regular.sum() >= 2
    rel = np.abs(hist[regular] - ref_hist[regular]) / ref_hist[regular]
    assert rel.max() <= REL_HIST, rel.max()
    noisy = ran & ~regular
    if noisy.any():
        assert hist[noisy].max() <= 1e-9 * h0
    if niters != ref_niters or not np.array_equal(np.isnan(hist), np.isnan(ref_hist)):
        k = min(niters, ref_niters)
        assert max(hist[k], ref_hist[k]) < 1e-140, (niters, ref_niters, hist[k], ref_hist[k])
        assert hist[niters] == 0.0 or ref_hist[ref_niters] == 0.0
    return rel.max()


def check_solution(x, ref_x=None):
    """Known answer x -> 1 (b = A*1, generate_matrix.cpp:284-286).  When the loop exits through exact underflow the
    reference's last iteration computes alpha = 0/0 and returns x = NaN (HPCCG.cpp:382-383); parity means NaN too."""
    x = np.asarray(x)
    if ref_x is not None and np.isnan(ref_x).any():
        assert np.isnan(x).all() == np.isnan(ref_x).all()
        return
    if np.isnan(x).any():
        assert ref_x is None or np.isnan(ref_x).any(), "NaN solution where the reference has none"
        return
    assert np.abs(x - 1.0).max() <= 1e-12


@pytest.mark.parametrize("dims,stencil", [((20, 30, 10), 27), ((20, 30, 10), 7), ((10, 10, 10), 27), ((33, 17, 5), 27),
                                           ((64, 64, 64), 27)])
def test_hpccg_matches_reference(H, refwrap, cuda, dims, stencil):
    H.set_rank(0, 1)
    H.set_options(stencil, True)
    A = H.generate_matrix(*dims)
    x = A.x.copy()
    niters, normr, times, hist = H.HPCCG(A, A.b, x, 150, 0.0)
    with refwrap.RefWorld(*dims, stencil=stencil, variant=ref_variant()) as R:
        ref = R.solve(150)
    check_history(hist, ref["hist"], niters, ref["niters"])
    assert normr == hist[niters]
    check_solution(x, ref["x"][0])
    assert times[0] > 0 and times[3] > 0
    A.destroy()


def test_noise_regime_relaxation_is_evidenced(H, refwrap, cuda, capsys):
    """The relaxed bar of check_history beyond ten decades is not an assertion of convenience: the reference's own serial
    and OpenMP builds agree to <= 1e-8 inside the regular regime and drift apart by DECADES beyond it, on the very problem
    where the first B200 run was 2.9 decades from the serial reference (10^3).  The spread is printed next to ours."""
    spread = reference_self_spread((10, 10, 10))
    if spread is None:
        pytest.skip("needs the serial and the OpenMP build of the reference (oracle/_ref)")
    first, rel, decades = spread
    H.set_rank(0, 1)
    H.set_options(27, True)
    A = H.generate_matrix(10, 10, 10)
    x = A.x.copy()
    niters, normr, _, hist = H.HPCCG(A, A.b, x, 150, 0.0)
    with refwrap.RefWorld(10, 10, 10, variant="serial") as R:
        ref = R.solve(150)["hist"]
    ran = ~np.isnan(hist) & ~np.isnan(ref) & (hist > 0) & (ref > 0)
    noisy = ran & (ref < 1e-10 * ref[0])
    ours = float(np.abs(np.log10(hist[noisy] / ref[noisy])).max()) if noisy.any() else 0.0
    with capsys.disabled():
        print(f"\n[noise regime, 10x10x10] starts at iteration {first}; reference serial vs OpenMP: {rel:.1e} relative before it, "
              f"{decades:.1f} decades apart after it; B200 vs serial reference after it: {ours:.1f} decades")
    assert rel <= 1e-8          # the 1e-8 bar is meaningful where the reference agrees with itself ...
    assert decades >= 0.5       # ... and it visibly does not beyond ten decades
    A.destroy()


def test_hpccg_against_golden_fixture(H, cuda):
    """Same check against the committed vectors generated from the real reference (tests/golden/make_golden.py)."""
    g = json.loads((GOLDEN / "golden.json").read_text())
    for rec in g["configs"]:
        if rec["ranks"] != 1 or rec["dims"] in ([1, 1, 1], [7, 1, 1]):
            continue
        H.set_rank(0, 1)
        H.set_options(rec["stencil"], True)
        A = H.generate_matrix(*rec["dims"])
        x = A.x.copy()
        niters, normr, _, hist = H.HPCCG(A, A.b, x, rec["max_iter"], 0.0)
        ref_hist = np.array([float.fromhex(v) for v in rec["hist"]])
        check_history(hist, ref_hist, niters, rec["niters"])
        A.destroy()


def test_unfused_sequence_agrees(H, cuda):
    """HPCCG_SOLVE_UNFUSED runs the reference's literal kernel sequence; both orders stay within the bar."""
    torch = cuda
    H.set_rank(0, 1)
    H.set_options(27, True)
    A = H.generate_matrix(20, 30, 10)
    m = A.device()
    b = torch.from_numpy(A.b.copy()).cuda()
    out = {}
    for name, flags in (("fused", 0), ("unfused", 1)):
        x = torch.zeros(A.local_nrow, dtype=torch.float64, device="cuda")
        out[name] = H.dev.cg_solve(m, b, x, 150, 0.0, flags=flags)
        assert (x.cpu().numpy() - 1.0).__abs__().max() <= 1e-12
    check_history(out["unfused"]["hist"], out["fused"]["hist"], out["unfused"]["niters"], out["fused"]["niters"])
    A.destroy()


def test_tolerance_exit_matches_reference(H, refwrap, cuda):
    """`normr > tolerance` (HPCCG.cpp:358) ends the loop at the same iteration as the reference."""
    H.set_rank(0, 1)
    H.set_options(27, True)
    A = H.generate_matrix(20, 30, 10)
    with refwrap.RefWorld(20, 30, 10, variant=ref_variant()) as R:
        for tol in (1e-3, 1e-9, 400.0, 1e300):
            ref = R.solve(150, tol)
            x = A.x.copy()
            niters, normr, _, hist = H.HPCCG(A, A.b, x, 150, tol)
            assert niters == ref["niters"], tol
            assert abs(normr - ref["normr"]) <= 1e-8 * ref["normr"]
    A.destroy()


@pytest.mark.parametrize("max_iter", [0, 1, 2, 3, 17])
def test_short_iteration_counts(H, refwrap, cuda, max_iter):
    """`for (k = 1; k < max_iter && normr > tolerance; k++)` (HPCCG.cpp:358): max_iter <= 1 runs no iteration, niters = 0,
    normr = the initial residual, x untouched; otherwise max_iter - 1 iterations."""
    H.set_rank(0, 1)
    H.set_options(27, True)
    A = H.generate_matrix(12, 9, 7)
    x = A.x.copy()
    niters, normr, _, hist = H.HPCCG(A, A.b, x, max_iter, 0.0)
    with refwrap.RefWorld(12, 9, 7, variant=ref_variant()) as R:
        ref = R.solve(max_iter)
    assert niters == ref["niters"] == max(max_iter - 1, 0)
    assert abs(normr - ref["normr"]) <= 1e-12 * ref["normr"]
    assert np.allclose(x, ref["x"][0], rtol=1e-12, atol=1e-14)
    A.destroy()


def test_solve_with_16_byte_aligned_vectors(H, refwrap, cuda):
    """b and x that are 16- but not 32-byte aligned take the 128-bit form of the vector kernels; same history bar."""
    torch = cuda
    H.set_rank(0, 1)
    H.set_options(27, True)
    A = H.generate_matrix(33, 17, 5)
    m = A.device()
    n = A.local_nrow
    bb = torch.zeros(n + 6, dtype=torch.float64, device="cuda")
    xx = torch.zeros(n + 6, dtype=torch.float64, device="cuda")
    off = 2 if bb.data_ptr() % 32 == 0 else 4  # make the slices 16- but not 32-byte aligned
    b, x = bb[off:off + n], xx[off:off + n]
    assert b.data_ptr() % 32 == 16 and x.data_ptr() % 32 == 16
    b.copy_(torch.from_numpy(A.b))
    out = H.dev.cg_solve(m, b, x, 150, 0.0)
    with refwrap.RefWorld(33, 17, 5, variant=ref_variant()) as R:
        ref = R.solve(150)
    check_history(out["hist"], ref["hist"], out["niters"], ref["niters"])
    check_solution(x.cpu().numpy(), ref["x"][0])
    assert xx[:off].abs().max().item() == 0.0 and xx[off + n:].abs().max().item() == 0.0  # nothing written outside x
    A.destroy()


@pytest.mark.parametrize("dims,stencil", [((20, 30, 10), 27), ((20, 30, 10), 7), ((10, 10, 10), 27), ((33, 17, 5), 27),
                                           ((16, 16, 16), 27), ((1, 1, 1), 27), ((7, 1, 1), 27), ((40, 8, 3), 7)])
def test_persistent_single_kernel_solve(H, refwrap, cuda, dims, stencil):
    """HPCCG_SOLVE_PERSISTENT (what HPCCG() uses for launch-bound sizes): the whole solve is ONE kernel of one thread-block
    cluster with the matrix blocks resident in shared memory and two cluster barriers per iteration.  Same bars as the
    normal loop; repeated solves are bit-identical; the tolerance exit takes the reference's iteration count."""
    torch = cuda
    H.set_rank(0, 1)
    H.set_options(stencil, True)
    A = H.generate_matrix(*dims)
    m = A.device()
    n = A.local_nrow
    b = torch.from_numpy(A.b.copy()).cuda()
    with refwrap.RefWorld(*dims, stencil=stencil, variant=ref_variant()) as R:
        ref = R.solve(150)
        rt = R.solve(150, 1e-6)
    launches = []
    outs = []
    for _ in range(2):
        xd = torch.zeros(n, dtype=torch.float64, device="cuda")
        l0 = H.launch_count()
        out = H.dev.cg_solve(m, b, xd, 150, 0.0, flags=H.SOLVE_PERSISTENT)
        launches.append(H.launch_count() - l0)
        check_history(out["hist"], ref["hist"], out["niters"], ref["niters"])
        check_solution(xd.cpu().numpy(), ref["x"][0])
        outs.append(out["hist"])
    assert np.array_equal(outs[0], outs[1], equal_nan=True)
    assert launches[1] == 1, launches  # one kernel for the whole solve
    xd = torch.zeros(n, dtype=torch.float64, device="cuda")
    out = H.dev.cg_solve(m, b, xd, 150, 1e-6, flags=H.SOLVE_PERSISTENT)
    assert out["niters"] == rt["niters"]
    if rt["normr"] >= 1e-10 * ref["hist"][0]:  # beyond that the recursion is rounding noise (check_history)
        assert abs(out["normr"] - rt["normr"]) <= 1e-8 * rt["normr"]
    out = H.dev.cg_solve(m, b, xd, 1, 0.0, flags=H.SOLVE_PERSISTENT)  # max_iter 1: set-up only (HPCCG.cpp:358)
    assert out["niters"] == 0
    # the reference-named call takes this path by itself below 2^20 rows
    x = A.x.copy()
    niters, normr, times, hist = H.HPCCG(A, A.b, x, 150, 0.0)
    assert np.array_equal(hist[:niters + 1], outs[0][:niters + 1], equal_nan=True) and times[0] > 0
    A.destroy()


def test_persistent_solve_declines_what_does_not_fit(H, cuda):
    """More rows than one cluster's shared memory holds: the flag is ignored and the normal loop runs (many launches)."""
    torch = cuda
    H.set_rank(0, 1)
    H.set_options(27, False)
    A = H.generate_matrix(48, 48, 48)
    m = A.device()
    b = torch.from_numpy(A.b.copy()).cuda()
    xd = torch.zeros(A.local_nrow, dtype=torch.float64, device="cuda")
    l0 = H.launch_count()
    out = H.dev.cg_solve(m, b, xd, 150, 0.0, flags=H.SOLVE_PERSISTENT)
    assert out["niters"] == 149 and H.launch_count() - l0 > 100
    assert (xd - 1.0).abs().max().item() <= 1e-12
    H.set_options(27, True)
    A.destroy()


def test_graph_replay_of_repeated_solves_is_bit_identical(H, refwrap, cuda, monkeypatch):
    """HPCCG_SOLVE_GRAPH (what HPCCG() uses below 2^20 rows when the single-kernel solve does not apply): the first solve
    with a key runs directly, the second is captured into a CUDA graph, later ones replay it.  Same kernels, same grids,
    same reduction trees -> same bits."""
    torch = cuda
    monkeypatch.setenv("HPCCG_B200_NO_PERSISTENT", "1")
    H.set_rank(0, 1)
    H.set_options(27, True)
    A = H.generate_matrix(20, 30, 10)
    with refwrap.RefWorld(20, 30, 10, variant=ref_variant()) as R:
        ref = R.solve(150)
    hists = []
    for _ in range(4):  # direct, capture + launch, replay, replay
        x = A.x.copy()
        niters, normr, times, hist = H.HPCCG(A, A.b, x, 150, 0.0)
        check_history(hist, ref["hist"], niters, ref["niters"])
        check_solution(x, ref["x"][0])
        assert times[0] > 0 and times[3] > 0
        hists.append(hist)
    for h in hists[1:]:
        assert np.array_equal(h, hists[0], equal_nan=True)
    # a different max_iter is a different key: direct again, then captured again
    for _ in range(3):
        x = A.x.copy()
        niters, normr, _, hist = H.HPCCG(A, A.b, x, 40, 0.0)
        assert niters == 39 and np.array_equal(hist[:40], hists[0][:40])
    # device pointers through the C-ABI, tolerance exit inside a replayed graph
    m = A.device()
    b = torch.from_numpy(A.b.copy()).cuda()
    xd = torch.zeros(A.local_nrow, dtype=torch.float64, device="cuda")
    with refwrap.RefWorld(20, 30, 10, variant=ref_variant()) as R:
        rt = R.solve(150, 1e-6)
    for _ in range(3):
        xd.zero_()
        out = H.dev.cg_solve(m, b, xd, 150, 1e-6, flags=H.SOLVE_GRAPH)
        assert out["niters"] == rt["niters"] and abs(out["normr"] - rt["normr"]) <= 1e-8 * rt["normr"]
    A.destroy()


def test_degenerate_sizes(H, refwrap, cuda):
    """1x1x1 converges exactly (normr underflows to 0 and the loop exits)."""
    for dims in ((1, 1, 1), (7, 1, 1), (2, 2, 1)):
        H.set_rank(0, 1)
        H.set_options(27, True)
        A = H.generate_matrix(*dims)
        x = A.x.copy()
        niters, normr, _, hist = H.HPCCG(A, A.b, x, 150, 0.0)
        with refwrap.RefWorld(*dims, variant=ref_variant()) as R:
            ref = R.solve(150)
        check_history(hist, ref["hist"], niters, ref["niters"])  # exit through exact underflow: count may differ
        assert np.isnan(x).all() and np.isnan(ref["x"][0]).all()  # alpha = 0/0 in the last iteration (HPCCG.cpp:382)
        A.destroy()


def _build_ranks(H, dims, size, stencil, host_arrays):
    def body(r):
        H.set_options(stencil, host_arrays)
        A = H.generate_matrix(*dims)
        H.make_local_matrix(A)
        return A
    return H.run_local_world(size, body)


@pytest.mark.parametrize("dims,size,stencil", [((16, 16, 8), 2, 27), ((4, 3, 2), 3, 27), ((12, 10, 1), 3, 27),
                                                ((16, 16, 8), 8, 27), ((5, 4, 3), 2, 7), ((32, 32, 16), 4, 27)])
def test_multi_rank_group_matches_mpi_reference(H, refwrap, cuda, dims, size, stencil):
    """N z-stacked ranks advanced in lock step on ONE GPU (halo by device copies, scalars summed in rank order)
    against the reference's -DUSING_MPI build run on thread ranks."""
    torch = cuda
    mats = _build_ranks(H, dims, size, stencil, True)
    ms = [A.device() for A in mats]
    bs = [torch.from_numpy(A.b.copy()).cuda() for A in mats]
    xs = [torch.zeros(A.local_nrow, dtype=torch.float64, device="cuda") for A in mats]
    out = H.dev.cg_solve_group(ms, bs, xs, 150, 0.0)
    with refwrap.RefWorld(*dims, size=size, stencil=stencil, variant=ref_variant(size)) as R:
        ref = R.solve(150)
    check_history(out["hist"], ref["hist"], out["niters"], ref["niters"])
    for x, rx in zip(xs, ref["x"]):
        check_solution(x.cpu().numpy(), rx)
    for A in mats:
        A.destroy()


def test_config2_split_over_four_ranks_matches_serial_golden(H, cuda):
    """Global 256^3 (BASELINE config #2) as 4 z-slabs of 256x256x64, device-generated, advanced in lock step on one GPU:
    the residual history must agree with the SERIAL reference's golden 256^3 history to the same 1e-8 (the decomposition
    only changes the reduction order; SURVEY.md section 4 verified this on the reference itself)."""
    torch = cuda
    gpath = GOLDEN / "golden_256.json"
    g = json.loads(gpath.read_text())
    mats = _build_ranks(H, (256, 256, 64), 4, 27, False)
    ms = [A.device() for A in mats]
    bs = [torch.from_numpy(A.b).cuda() for A in mats]
    xs = [torch.zeros(A.local_nrow, dtype=torch.float64, device="cuda") for A in mats]
    out = H.dev.cg_solve_group(ms, bs, xs, 150, 0.0)
    ref_hist = np.array([float.fromhex(v) for v in g["hist"]])
    worst = check_history(out["hist"], ref_hist, out["niters"], g["niters"])
    print("4-rank 256^3 worst relative residual difference vs serial reference:", worst)
    for x in xs:
        assert (x - 1.0).abs().max().item() <= 1e-12
    for A in mats:
        A.destroy()


@pytest.mark.parametrize("dims,size,stencil", [((20, 30, 10), 1, 27), ((20, 30, 10), 1, 7), ((16, 16, 8), 3, 27),
                                                ((12, 10, 1), 3, 27), ((5, 4, 3), 2, 7), ((3, 1, 2), 2, 27)])
def test_device_generated_ell_is_bit_identical(H, cuda, dims, size, stencil):
    """generate_matrix with host_arrays=0 builds the ELL mirror on the device; it must equal the mirror
    repacked from the host rows (which are bit-exact with the reference, tests/test_host_setup.py)."""
    host = _build_ranks(H, dims, size, stencil, True) if size > 1 else None
    devo = _build_ranks(H, dims, size, stencil, False) if size > 1 else None
    if size == 1:
        H.set_rank(0, 1)
        H.set_options(stencil, True)
        host = [H.generate_matrix(*dims)]
        H.set_options(stencil, False)
        devo = [H.generate_matrix(*dims)]
        H.set_options(stencil, True)
    for Ah, Ad in zip(host, devo):
        vh, ch = Ah.device().download()
        vd, cd = Ad.device().download()
        assert Ah.device().info() == Ad.device().info()
        assert np.array_equal(ch, cd) and np.array_equal(vh, vd)
        assert np.array_equal(Ah.b, Ad.b) and np.array_equal(Ah.x, Ad.x) and np.array_equal(Ah.xexact, Ad.xexact)
        for f in ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length"):
            assert np.array_equal(Ah.array(f), Ad.array(f)), f
        Ah.destroy()
        Ad.destroy()


def test_full_size_config2_properties(H, cuda):
    """BASELINE config #2 (256^3, 27-pt, 150 iterations) with the device-side generator: size-independent
    properties (A*1 == b, x -> 1) plus the golden residuals of the real reference."""
    torch = cuda
    H.set_rank(0, 1)
    H.set_options(27, False)
    A = H.generate_matrix(256, 256, 256)
    H.set_options(27, True)
    m = A.device()
    n = A.local_nrow
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    H.dev.spmv(m, ones, y)
    b = torch.from_numpy(A.b).cuda()
    assert torch.equal(y, b)  # row sums: 27 - (nnz-1), exact in fp64
    # 27*n - sum(b) counts the off-diagonal entries: closed-form nnz (3nx-2)(3ny-2)(3nz-2)
    assert int(round((27.0 * n - b.sum().item()))) + n == (3 * 256 - 2) ** 3
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    out = H.dev.cg_solve(m, b, x, 150, 0.0)
    assert out["niters"] == 149
    assert (x - 1.0).abs().max().item() <= 1e-12
    gpath = GOLDEN / "golden_256.json"
    if gpath.exists():
        g = json.loads(gpath.read_text())
        ref_hist = np.array([float.fromhex(v) for v in g["hist"]])
        worst = check_history(out["hist"], ref_hist, out["niters"], g["niters"])
        print("256^3 worst relative residual difference vs serial reference:", worst)
    else:  # survey-time values of the serial reference
        assert abs(out["hist"][0] - 7475.3028032314514) <= 1e-8 * 7475.3028032314514
        assert abs(out["normr"] - 2.2419957139761042e-18) <= 1e-8 * 2.2419957139761042e-18
    A.destroy()


def test_compress_invalidates_a_captured_solve(H, refwrap, cuda):
    """A solve captured as a CUDA graph refers to the SELL arrays; hpccg_dev_matrix_compress releases them, so the graph
    must go with them (a replay would run the old kernels on freed memory).  After the format switch the same call
    sequence still gives the reference's history."""
    torch = cuda
    H.set_rank(0, 1)
    H.set_options(27, False)
    A = H.generate_matrix(48, 48, 24)  # too many rows for the single-kernel solve: graph path
    m = A.device()
    n = A.local_nrow
    b = torch.from_numpy(A.b.copy()).cuda()
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    hists = []
    for _ in range(3):  # direct, capture + launch, replay
        x.zero_()
        hists.append(H.dev.cg_solve(m, b, x, 60, 0.0, flags=H.SOLVE_GRAPH)["hist"])
    assert np.array_equal(hists[0], hists[2], equal_nan=True)
    assert m.compress()["format"] == 1
    for _ in range(3):  # same key as before the switch
        x.zero_()
        out = H.dev.cg_solve(m, b, x, 60, 0.0, flags=H.SOLVE_GRAPH)
        assert out["niters"] == 59
        rel = np.abs(out["hist"][:60] - hists[0][:60]) / hists[0][:60]
        assert rel.max() <= 1e-8
    H.set_options(27, True)
    A.destroy()


def test_localised_matrix_is_not_solved_as_a_single_rank(H, cuda):
    """A matrix with halo columns solved under a 1-rank context would read a halo nobody fills: refused, not wrong."""
    torch = cuda
    mats = _build_ranks(H, (8, 8, 4), 2, 27, True)
    H.set_rank(0, 1)
    m = mats[0].device()
    n = mats[0].local_nrow
    b = torch.from_numpy(mats[0].b.copy()).cuda()
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    with pytest.raises(H.HpccgError, match="halo columns"):
        H.dev.cg_solve(m, b, x, 10, 0.0)
    for A in mats:
        A.destroy()


def test_second_device_gets_its_own_kernel_setup(H, refwrap, cuda):
    """Occupancy results and the dynamic-shared-memory opt-in of the TMA kernels are per device: a process that moves to
    a second GPU (hpccg_set_device) must be able to run the 81-166 KB kernels there."""
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from hpccg_sycl_b200._capi import lib, check
    H.set_rank(0, 1)
    H.set_options(27, True)
    with refwrap.RefWorld(20, 30, 10, variant=ref_variant()) as R:
        ref = R.solve(150)
    try:
        for dev in (0, 1):
            check(lib.hpccg_set_device(dev))
            torch.cuda.set_device(dev)
            A = H.generate_matrix(20, 30, 10)
            m = A.device()
            b = torch.from_numpy(A.b.copy()).to(f"cuda:{dev}")
            x = torch.zeros(A.local_nrow, dtype=torch.float64, device=f"cuda:{dev}")
            out = H.dev.cg_solve(m, b, x, 150, 0.0)  # normal loop: TMA SpMV
            check_history(out["hist"], ref["hist"], out["niters"], ref["niters"])
            A.destroy()
    finally:
        check(lib.hpccg_set_device(0))
        torch.cuda.set_device(0)


def test_pageable_and_page_locked_host_vectors_give_the_same_solve(H, cuda):
    """HPCCG() with the caller's own pageable vectors (what the reference's `new double[]` are): large transfers go through the
    library's page-locked bounce buffers in 64 MB chunks (84 MB here: two chunks, the second one partial); same bits as with
    the page-locked vectors generate_matrix hands out."""
    H.set_rank(0, 1)
    H.set_options(27, False)
    A = H.generate_matrix(256, 256, 160)
    n = A.local_nrow
    x1 = A.x  # page-locked (generate_matrix registers what it hands out)
    x1[:] = 0.0
    it1, nr1, _, h1 = H.HPCCG(A, A.b, x1, 12, 0.0)
    bp = np.array(A.b, copy=True)  # pageable copies
    xp = np.zeros(n)
    it2, nr2, _, h2 = H.HPCCG(A, bp, xp, 12, 0.0)
    assert it1 == it2 == 11 and nr1 == nr2
    assert np.array_equal(h1, h2, equal_nan=True)
    assert np.array_equal(x1, xp)
    H.set_options(27, True)
    A.destroy()
