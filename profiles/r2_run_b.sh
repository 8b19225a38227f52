#!/usr/bin/env bash
# pattern SpMV A/B: classic (one gather per entry) vs z-marching kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pattern.py tests/test_gpu_kernels.py -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
for mode in classic march; do
  HPCCG_B200_PATTERN=$mode timeout 600 python bench.py --format pattern --no-also --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2b_pattern_$mode.json 2> gpurun_out/r2b_pattern_$mode.err; echo "$mode rc=$?"
  tail -c 300 gpurun_out/r2b_pattern_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_pattern_$mode.json").read().strip().splitlines()[-1])
print("$mode", d["value"], d["roofline"]["kernels"], d["check"])
PY
done
HPCCG_B200_PATTERN=march timeout 600 python bench.py --format pattern --workload c3 --no-also --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/r2b_pattern_c3.json 2>&1; echo "c3 rc=$?"
HPCCG_B200_PATTERN=march timeout 600 python bench.py --format pattern --workload c2 --no-also --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/r2b_pattern_c2.json 2>&1; echo "c2 rc=$?"
