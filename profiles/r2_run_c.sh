#!/usr/bin/env bash
# ncu --set full of the z-marching pattern kernel (512^3), after a plain run
mkdir -p gpurun_out
CMD="python bench.py --format pattern --no-also --no-cpu-baseline --no-e2e --steps 1 --warmup 1"
$CMD > gpurun_out/r2c_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r2c_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k 'regex:spmv_pattern_march_kernel' -s 20 -c 1 -f -o gpurun_out/r2c_march $CMD > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/r2c_march.ncu-rep
