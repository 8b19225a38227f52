#!/usr/bin/env bash
# 8 GPUs: the torchrun parity worker at 2 / 4 / 8 ranks (peer-memory and NCCL planes, pattern format, thin blocks), then the
# default bench line at N = 8 (weak 512^3 per GPU + parity block + strong 512x512x1024) and at N = 4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2k_topo.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2k_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest_multi.log
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err; echo "bench n$N rc=$?"
  tail -c 300 gpurun_out/r2k_bench_n$N.err
done
