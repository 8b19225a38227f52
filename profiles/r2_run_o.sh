#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solve.py tests/test_gpu_multi.py tests/test_gpu_cli.py -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2o_bench_n2.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["parity"]["ok"], d["parity"]["worst_rel"], round(d["also_strong"]["value"],1))
PY
