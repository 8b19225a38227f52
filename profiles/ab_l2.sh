#!/usr/bin/env bash
# profiles/ab_l2.sh <tag> -- A/B of the TMA SpMV's L2 prefetch distance on the default workload (27-pt 512^3).
TAG="${1:-abl2}"
B="python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline --no-e2e"
for v in 0 2 4 8; do
  HPCCG_B200_L2_AHEAD=$v timeout 300 $B 2>&1 | tail -1 > gpurun_out/${TAG}_l2ahead$v.json
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d=json.loads(open(f).read()); k=d["roofline"]["kernels"]
        print(f, "GF/s %.1f"%d["value"], "spmv %.3f ms %.0f GB/s"%(k["spmv_dot"]["ms"],k["spmv_dot"]["gbs"]), "iter %.3f ms"%k["iteration"]["ms"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "FAILED", e, open(f).read()[-300:])
PY
