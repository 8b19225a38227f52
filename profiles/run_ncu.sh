#!/usr/bin/env bash
# profiles/run_ncu.sh -- the ncu recipe of B200_PROFILING.md for this repo (run under gpurun, 1 GPU).
# Usage: profiles/run_ncu.sh <tag> [bench flags]  -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_full.ncu-rep
#   e.g. profiles/run_ncu.sh r02_c3 --workload c3      profiles/run_ncu.sh r02_pattern --format pattern
set -uo pipefail
TAG="${1:-r02}"
shift || true
CMD="python bench.py --steps 1 --warmup 1 --no-also --no-cpu-baseline --no-e2e $*"
mkdir -p gpurun_out
# plain run first: a command is profiled only after it has exited 0 without ncu
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
if [ "${NCU_LAUNCH_LIST:-1}" = "1" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
      $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
  echo "launch list rc=$?"
fi
# the loop's kernels, once per change: fused SpMV+dot, r-update+dot, p/x-update (skip the first solve's launches)
ncu --set full --clock-control none --import-source on \
    -k 'regex:spmv_sell_tma_kernel|spmv_ell_kernel|spmv_pattern_march_kernel|spmv_pattern_kernel|vec_stream_tma_kernel|update_r_dot_kernel|p_update_x_kernel' \
    -s 30 -c 6 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
