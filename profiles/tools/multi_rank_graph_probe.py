"""torchrun worker: launch-bound multi-rank solve (64x64x16 per rank) with direct launches vs CUDA-graph replay on the peer plane."""
import json, os, statistics, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
local_rank = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
import hpccg_pkg
H = hpccg_pkg.load()
from hpccg_sycl_b200 import dist as hdist
rank, size, _ = hdist.init_process_group_context(use_nccl=True)
H.set_print(False); H.set_options(27, True)
out = {}
for dims in ((64, 64, 16), (128, 128, 32)):
    A = H.generate_matrix(*dims); H.make_local_matrix(A); m = A.device(); n = A.local_nrow
    b = torch.from_numpy(A.b.copy()).cuda(); x = torch.zeros(n, dtype=torch.float64, device="cuda")
    res = {}
    for name, fl in (("direct", 0), ("graph", 16)):
        ms = []
        for i in range(24):
            x.zero_(); dist.barrier(device_ids=[local_rank]); torch.cuda.synchronize(); t0 = time.perf_counter()
            o = H.dev.cg_solve(m, b, x, 150, 0.0, flags=fl, want_hist=False)
            ms.append((time.perf_counter() - t0) * 1e3)
        res[name] = {"call_ms": statistics.median(ms[4:]), "niters": o["niters"], "normr": o["normr"], "x_err": float((x - 1).abs().max())}
    out["x".join(map(str, dims))] = res
    A.destroy()
hdist.finalize()
if rank == 0:
    print(json.dumps({"ranks": size, "comm": "peer memory", "solves": out}))
dist.destroy_process_group()
