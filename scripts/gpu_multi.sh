#!/bin/bash
# 2-GPU run: all gpu tests (incl. NCCL merge + C++ host driver) and bench at N=1,2
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi.log
tail -6 gpurun_out/test_multi.log
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 2500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
