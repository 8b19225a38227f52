#!/usr/bin/env bash
# final 1-GPU sanity of the round: smoke(), full GPU suite, default bench line and the reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2m_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_reference.json 2> gpurun_out/r2m_bench_reference.err; echo "reference rc=$?"
