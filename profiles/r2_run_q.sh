#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pattern.py -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
for mode in 1 2; do
  HPCCG_B200_MARCH=$mode timeout 600 python bench.py --format pattern --no-also --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/r2q_march$mode.json 2> gpurun_out/r2q_march$mode.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_march$mode.json").read().strip().splitlines()[-1])
print("lines per thread=$mode", round(d["value"],1), d["roofline"]["kernels"]["spmv_dot"], d["check"])
PY
done
