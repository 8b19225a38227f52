#!/usr/bin/env bash
# ncu evidence of the round-2 default path and of the other configs' dominant kernels; raw metric pages are extracted on the
# box (the reports themselves are too large to bring back together)
bash profiles/run_ncu.sh r02_weak512
NCU_LAUNCH_LIST=0 bash profiles/run_ncu.sh r02_c2 --workload c2
NCU_LAUNCH_LIST=0 bash profiles/run_ncu.sh r02_c3 --workload c3
NCU_LAUNCH_LIST=0 bash profiles/run_ncu.sh r02_pattern --format pattern
for t in r02_weak512 r02_c2 r02_c3 r02_pattern; do
  ncu -i gpurun_out/${t}_full.ncu-rep --page raw --csv > gpurun_out/${t}_raw.csv 2>/dev/null
  [ "$t" != "r02_weak512" ] && rm -f gpurun_out/${t}_full.ncu-rep
done
ls -la gpurun_out | grep r02_
