#!/usr/bin/env bash
# one 8-GPU lease: N = 1 and N = 8 lines back to back (same node, same clocks regime) for the same-lease weak / strong efficiencies
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_n1.json 2> gpurun_out/r2v_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2v_n8.json 2> gpurun_out/r2v_n8.err; echo "n8 rc=$?"
python - <<'PY'
import json
a=json.loads(open("gpurun_out/r2v_n1.json").read().strip().splitlines()[-1]); b=json.loads(open("gpurun_out/r2v_n8.json").read().strip().splitlines()[-1])
print("weak", round(a["value"],1), round(b["value"],1), "eff", round(b["value"]/8/a["value"],4), "e2e eff", round(b["e2e"]["value"]/8/a["e2e"]["value"],4))
print("strong", round(a["also_strong"]["value"],1), round(b["also_strong"]["value"],1), "eff", round(b["also_strong"]["value"]/8/a["also_strong"]["value"],4))
print(a["clocks"], b["clocks"])
PY
