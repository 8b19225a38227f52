"""Worker of tests/test_gpu_multi.py: one rank (one GPU) of a torchrun job.  The N-rank CG solve through the
reference-named API (NCCL halo send/recv overlapped with the interior SpMV + NCCL scalar gathers) is compared with
the oracle's N-rank world (the reference's -DUSING_MPI build on thread ranks where oracle/_ref has it)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    out_dir = Path(sys.argv[1])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import hpccg_pkg
    H = hpccg_pkg.load()
    from hpccg_sycl_b200 import dist as hdist
    import refwrap
    from test_gpu_solve import check_history, check_solution
    rank, size, _ = hdist.init_process_group_context(use_nccl=True)
    H.set_print(False)
    variant = "mpi" if refwrap.available("mpi") else "oracle"
    result = {"rank": rank, "size": size, "cases": []}
    cases = [((16, 16, 8), 27, True), ((5, 4, 3), 7, True), ((12, 10, 1), 27, True), ((64, 64, 16), 27, True),
             ((64, 64, 16), 27, False), ((96, 80, 24), 27, False), ((64, 64, 16), 7, False)]
    # every case twice: halos + scalar sums through peer memory inside the kernels (default), then the NCCL path
    runs = [(c, "sell", k) for c in ("p2p", "nccl") for k in cases]
    # the opt-in pattern-coded mirror through both communication paths
    runs += [(c, "pattern", k) for c in ("p2p", "nccl") for k in (cases[0], cases[4], cases[6])]
    for comm, fmt, (dims, stencil, host_rows) in runs:
        os.environ["HPCCG_B200_COMM"] = comm  # read when the matrix's peer link is created (first solve)
        H.set_matrix_format(fmt)
        H.set_options(stencil, host_rows)
        A = H.generate_matrix(*dims)
        H.make_local_matrix(A)
        n = A.local_nrow
        with refwrap.RefWorld(*dims, size=size, stencil=stencil, variant=variant) as R:
            ref = R.solve(150)
            # exchange_externals + HPC_sparsemv + ddot on seeded vectors, every rank its own seed
            ncol = [R.scalar(r, "local_ncol") for r in range(size)]
            nrow = [R.scalar(r, "local_nrow") for r in range(size)]
            xs = [np.concatenate([np.random.default_rng(12345 + r).uniform(-1, 1, nrow[r]), np.zeros(ncol[r] - nrow[r])])
                  for r in range(size)]
            ys = R.spmv([v.copy() for v in xs], exchange=True)
            dref = R.ddot([v[:nrow[r]].copy() for r, v in enumerate(xs)], ys)[0]
        for flags_env in ("0", "1"):
            os.environ["HPCCG_B200_UNFUSED"] = flags_env
            x = A.x.copy()
            niters, normr, times, hist = H.HPCCG(A, A.b, x, 150, 0.0)
            worst = check_history(hist, ref["hist"], niters, ref["niters"])
            check_solution(x, ref["x"][rank])
        os.environ["HPCCG_B200_UNFUSED"] = "0"
        # repeated identical solves: below 2^20 rows HPCCG() captures the second one into a CUDA graph and replays it -- on the
        # peer-memory plane that works for one rank of a multi-GPU job as well (the NCCL plane keeps launching directly)
        first = None
        for rep in range(4):
            x = A.x.copy()
            nit, nr, _, hist = H.HPCCG(A, A.b, x, 150, 0.0)
            check_history(hist, ref["hist"], nit, ref["niters"])
            check_solution(x, ref["x"][rank])
            if first is None:
                first = hist
            assert np.array_equal(hist, first, equal_nan=True), (comm, fmt, dims, rep)
        # `normr > tolerance` ends the loop at the same iteration on every rank (HPCCG.cpp:358); the kernels enqueued after
        # that return at once on all ranks alike, so nobody is left waiting for a peer
        with refwrap.RefWorld(*dims, size=size, stencil=stencil, variant=variant) as R:
            rt = R.solve(150, 1e-3)
        xt = A.x.copy()
        nit, nr, _, _ = H.HPCCG(A, A.b, xt, 150, 1e-3)
        assert nit == rt["niters"] and abs(nr - rt["normr"]) <= 1e-8 * rt["normr"], (nit, rt["niters"], nr, rt["normr"])
        mine = xs[rank].copy()
        H.exchange_externals(A, mine)
        y = np.empty(n)
        H.HPC_sparsemv(A, mine, y)
        assert np.array_equal(y, ys[rank]), "HPC_sparsemv after exchange_externals differs from the reference"
        d, t_all = H.ddot(n, mine[:n].copy(), y)
        scale = sum(float(np.abs(xs[r][:nrow[r]] * ys[r]).sum()) for r in range(size))
        assert abs(d - dref) <= 1e-12 * scale, (d, dref)
        res = H.compute_residual(n, x, A.xexact)
        result["cases"].append({"dims": dims, "stencil": stencil, "host_rows": host_rows, "niters": niters, "comm": comm, "format": fmt,
                                "worst_rel": float(worst), "residual": float(res), "ddot_rel": abs(d - dref) / scale})
        A.destroy()
    hdist.finalize()
    (out_dir / f"rank{rank}.json").write_text(json.dumps(result))
    dist.barrier(device_ids=[local_rank])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
