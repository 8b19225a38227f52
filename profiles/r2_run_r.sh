#!/usr/bin/env bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:spmv_pattern_march2_kernel' -s 20 -c 1 -f -o gpurun_out/r2r_march2 python bench.py --format pattern --no-also --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r2r_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r2r_march2.ncu-rep --page raw --csv > gpurun_out/r2r_raw.csv 2>/dev/null
ncu -i gpurun_out/r2r_march2.ncu-rep --page source --csv > gpurun_out/r2r_src.csv 2>/dev/null
rm -f gpurun_out/r2r_march2.ncu-rep
