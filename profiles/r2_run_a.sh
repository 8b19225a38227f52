#!/usr/bin/env bash
# round-2 first GPU pass (2 GPUs): GPU test suite incl. the 2-rank torchrun worker, bench at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench n1 rc=$?"
tail -c 600 gpurun_out/r2a_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/r2a_bench_n2.err
cut -c1-600 gpurun_out/r2a_bench_n2.json
