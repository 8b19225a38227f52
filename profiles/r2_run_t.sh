#!/usr/bin/env bash
# 8 GPUs, final code of the round: parity worker at 8 ranks + the default line at N = 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "8" > gpurun_out/r2t_pytest_multi8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest_multi8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2t_bench_n8.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]; s=d["also_strong"]
print("weak", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {a:(round(k[a]["ms"],4) if isinstance(k[a],dict) else k[a]) for a in k})
print("strong", round(s["value"],1), round(s["ms_per_iteration"],4), {a:(round(s["kernels"][a]["ms"],4) if isinstance(s["kernels"][a],dict) else s["kernels"][a]) for a in s["kernels"]})
print("parity", d["parity"]["ok"], d["parity"]["worst_rel"])
PY
