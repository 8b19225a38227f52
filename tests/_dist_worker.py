"""Worker of tests/test_dist_gloo.py: one rank of a torch.distributed (gloo) job on CPU.  Exercises the N>1
host path that bench.py uses on GPUs: rank context from the process group, the set-up allgather installed on
the library, generate_matrix + make_local_matrix per rank, compared with the oracle's rank of the same world."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def main():
    out_dir = Path(sys.argv[1])
    dist.init_process_group("gloo")
    import hpccg_pkg
    H = hpccg_pkg.load()
    from hpccg_sycl_b200 import dist as hdist
    import refwrap
    rank, size, _ = hdist.init_process_group_context(use_nccl=False)
    assert H.get_rank() == (rank, size)
    result = {"rank": rank, "size": size, "cases": []}
    for dims, stencil in (((6, 5, 2), 27), ((8, 8, 4), 7), ((16, 12, 1), 27)):
        H.set_options(stencil, True)
        A = H.generate_matrix(*dims)
        H.make_local_matrix(A)
        variant = "mpi" if refwrap.available("mpi") else "oracle"
        ok = True
        with refwrap.RefWorld(*dims, size=size, stencil=stencil, variant=variant) as R:
            for s in ("start_row", "stop_row", "total_nrow", "local_nrow", "local_ncol", "num_external",
                      "num_send_neighbors", "total_to_be_sent"):
                ok = ok and A.scalar(s) == R.scalar(rank, s)
            for a in ("list_of_inds", "external_index", "external_local_index", "elements_to_send", "neighbors",
                      "recv_length", "send_length"):
                ok = ok and np.array_equal(A.array(a), R.array(rank, a))
        result["cases"].append({"dims": dims, "stencil": stencil, "ok": bool(ok), "ncol": A.local_ncol})
        A.destroy()
    # the bench's work split: every rank owns one z-slab; totals reduce across ranks
    import torch
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    result["max_reduce"] = t.item()
    (out_dir / f"rank{rank}.json").write_text(json.dumps(result))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
