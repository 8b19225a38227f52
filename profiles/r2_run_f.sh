#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solve.py -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2f_pytest.log
python - <<'PY' > gpurun_out/r2f_c1.json 2> gpurun_out/r2f_c1.err
import json, statistics, time, sys
sys.path.insert(0, ".")
import numpy as np, torch, hpccg_pkg
H = hpccg_pkg.load(); H.set_print(False); H.set_rank(0, 1)
res = {}
for dims in ((20, 30, 10), (16, 16, 16), (21, 21, 21)):
    H.set_options(27, True)
    A = H.generate_matrix(*dims); m = A.device(); n = A.local_nrow
    b = torch.from_numpy(A.b.copy()).cuda(); x = torch.zeros(n, dtype=torch.float64, device="cuda")
    r = {}
    for name, fl in (("graph", 16), ("persistent", 64)):
        dev, wall = [], []
        for i in range(43):
            x.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter()
            o = H.dev.cg_solve(m, b, x, 150, 0.0, flags=fl, want_hist=False)
            t1 = time.perf_counter()
            if i >= 3: dev.append(o["loop_ms"]); wall.append((t1 - t0) * 1e3)
        r[name] = {"device_ms": statistics.median(dev), "call_ms": statistics.median(wall), "niters": o["niters"], "normr": o["normr"],
                   "x_err": float((x - 1).abs().max())}
    res["x".join(map(str, dims))] = r
    A.destroy()
print(json.dumps(res))
PY
cat gpurun_out/r2f_c1.json; tail -3 gpurun_out/r2f_c1.err
