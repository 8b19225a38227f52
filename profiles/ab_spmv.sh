#!/usr/bin/env bash
# profiles/ab_spmv.sh <tag> -- A/B of the SpMV variants on the default workload (27-pt 512^3), one bench line each.
TAG="${1:-ab}"
B="python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline --no-e2e"
for v in 0 1 2; do
  HPCCG_B200_TMA_VARIANT=$v timeout 300 $B 2>&1 | tail -1 > gpurun_out/${TAG}_tma$v.json
done
HPCCG_B200_SPMV=reg timeout 300 $B 2>&1 | tail -1 > gpurun_out/${TAG}_reg.json
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d=json.loads(open(f).read()); k=d["roofline"]["kernels"]
        print(f, "GF/s %.1f"%d["value"], "spmv %.3f ms %.0f GB/s"%(k["spmv_dot"]["ms"],k["spmv_dot"]["gbs"]), "iter %.3f ms"%k["iteration"]["ms"], d["check"])
    except Exception as e: print(f, "FAILED", e, open(f).read()[-300:])
PY
