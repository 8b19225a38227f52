#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
for nb in 1 0; do
  if [ $nb = 1 ]; then export HPCCG_B200_NO_BOUNCE=1; else unset HPCCG_B200_NO_BOUNCE; fi
  timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-strong > gpurun_out/r2n_bench_nobounce$nb.json 2> gpurun_out/r2n_bench_nobounce$nb.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_bench_nobounce$nb.json").read().strip().splitlines()[-1])
print("NO_BOUNCE=$nb value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pageable", d["e2e"]["pageable"])
PY
done
