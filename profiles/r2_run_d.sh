#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pattern.py -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2d_pytest.log
CMD="python bench.py --format pattern --no-also --no-cpu-baseline --no-e2e --steps 3 --warmup 2"
$CMD > gpurun_out/r2d_march.json 2> gpurun_out/r2d_march.err; echo "march rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2d_march.json").read().strip().splitlines()[-1])
print("march", d["value"], d["roofline"]["kernels"]["spmv_dot"], d["check"])
PY
ncu --set full --clock-control none --import-source on -k 'regex:spmv_pattern_march_kernel' -s 20 -c 1 -f -o gpurun_out/r2d_march python bench.py --format pattern --no-also --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r2d_ncu.log 2>&1
echo "ncu rc=$?"
