#!/usr/bin/env bash
mkdir -p gpurun_out
cat > /tmp/persist_probe.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch, hpccg_pkg
H = hpccg_pkg.load(); H.set_print(False); H.set_rank(0, 1); H.set_options(27, True)
A = H.generate_matrix(20, 30, 10); m = A.device(); n = A.local_nrow
b = torch.from_numpy(A.b.copy()).cuda(); x = torch.zeros(n, dtype=torch.float64, device="cuda")
for i in range(4):
    x.zero_()
    o = H.dev.cg_solve(m, b, x, 150, 0.0, flags=64, want_hist=False)
print(o["loop_ms"], o["niters"])
PY
python /tmp/persist_probe.py
ncu --set full --clock-control none --import-source on -k 'regex:cg_cluster_kernel' -s 2 -c 1 -f -o gpurun_out/r2g_persist python /tmp/persist_probe.py > gpurun_out/r2g_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2g_ncu.log
