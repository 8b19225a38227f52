#!/usr/bin/env bash
# 2 GPUs: full GPU suite (incl. the 2-rank torchrun worker: peer / NCCL planes, pattern format, thin ragged blocks), bench at N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2i_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench n2 rc=$?"
tail -c 300 gpurun_out/r2i_bench_n2.err
