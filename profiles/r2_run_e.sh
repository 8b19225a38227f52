#!/usr/bin/env bash
# A/B: the loop's vector kernels with plain 256-bit LDG/STG ("0") vs bulk-async (TMA) streaming, tile (doubles) x stages
mkdir -p gpurun_out
HPCCG_B200_VEC_TMA=2048,3 timeout 600 python -m pytest tests/test_gpu_solve.py -x -q -k "matches_reference or config2" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
for mode in 0 1024,2 1024,3 1024,4 2048,2 2048,3 4096,2; do
  HPCCG_B200_VEC_TMA=$mode timeout 600 python bench.py --no-also --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/r2e_vectma_$mode.json 2> gpurun_out/r2e_vectma_$mode.err; echo "mode $mode rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2e_vectma_$mode.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("VEC_TMA=$mode", round(d["value"],1), {a:(round(k[a]["ms"],4), round(k[a]["gbs"])) for a in k}, d["check"]["normr"], d["check"]["x_max_err"], d["clocks"]["sm_mhz"])
PY
done
