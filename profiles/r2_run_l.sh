#!/usr/bin/env bash
# 2 GPUs: new regression tests + A/B of the put-first tile order in the bulk-async p-update (strong config, 512x512x1024 over 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_multi.py -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2l_pytest.log
for rotate in 0 1; do
  HPCCG_B200_PUT_ROTATE=$rotate timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$rotate bench.py --gpus 2 --workload strong --nz 64 --steps 10 --warmup 5 --no-e2e > gpurun_out/r2l_rotate$rotate.json 2> gpurun_out/r2l_rotate$rotate.err; echo "rotate $rotate rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2l_rotate$rotate.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("rotate=$rotate", round(d["value"],1), {a:(round(k[a]["ms"],4) if isinstance(k[a],dict) else k[a]) for a in k}, d["check"])
PY
done
