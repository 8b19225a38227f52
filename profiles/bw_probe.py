#!/usr/bin/env python
"""profiles/bw_probe.py -- practical HBM ceilings on this B200 for the access mixes of the CG kernels, measured with
this library's own streaming kernels (CUDA events, 20 launches after 5 warm-up, vectors >> L2):
  read-only 1 stream (ddot x.x), read-only 2 streams (ddot x.y), 2R+1W (waxpby), 4R+2W (update_xr_dot), torch copy 1R+1W.
Prints one JSON object; the numbers put the SpMV's achieved GB/s in context (is ~6.3 TB/s the DRAM ceiling?)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import hpccg_pkg  # noqa: E402

H = hpccg_pkg.load()
torch.cuda.set_device(0)
n = 1 << 28  # 2 GiB per vector
vs = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(4)]
out = torch.zeros(4, dtype=torch.float64, device="cuda")


def timeit(fn, nbytes, reps=20):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"ms": ms, "GBps": nbytes / ms / 1e6}


res = {
    "n_doubles": n,
    "read1_ddot_xx": timeit(lambda: H.dev.dot(n, vs[0], vs[0], out), 8 * n),
    "read2_ddot_xy": timeit(lambda: H.dev.dot(n, vs[0], vs[1], out), 16 * n),
    "r2w1_waxpby": timeit(lambda: H.dev.waxpby(n, 1.0, vs[0], 0.5, vs[1], vs[2]), 24 * n),
    "r4w2_update_xr_dot": timeit(lambda: H.dev.update_xr_dot(n, out.data_ptr() + 8, vs[0], vs[1], vs[2], vs[3], out), 48 * n),
    "r1w1_torch_copy": timeit(lambda: vs[2].copy_(vs[0]), 16 * n),
    "w1_torch_fill": timeit(lambda: vs[2].fill_(1.0), 8 * n),
}
print(json.dumps(res))
