#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2h_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2h_bench_n1.err
