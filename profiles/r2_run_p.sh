#!/usr/bin/env bash
# roofline point of the ragged SELL-C-sigma kernel: the 27-pt 256^3 matrix forced into format 2 (sigma 8192 and sigma 1) vs the TMA path
mkdir -p gpurun_out
for mode in tma ragged ragged_s1; do
  unset HPCCG_B200_RAGGED HPCCG_B200_SIGMA
  [ $mode = ragged ] && export HPCCG_B200_RAGGED=1
  [ $mode = ragged_s1 ] && export HPCCG_B200_RAGGED=1 HPCCG_B200_SIGMA=1
  timeout 600 python bench.py --workload c2 --no-also --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/r2p_$mode.json 2> gpurun_out/r2p_$mode.err; echo "$mode rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_$mode.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("$mode", round(d["value"],1), {a:(round(k[a]["ms"],4), round(k[a]["gbs"])) for a in k}, d["check"])
PY
done
