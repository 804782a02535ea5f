#!/bin/bash
# round 2, multi-GPU call: N-rank parity (pytest 2-GPU test + bench parity leg) and the bench line at N = $1 (default 2)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_n$N.csv 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_round2.py -m gpu -q -x --tb=short > gpurun_out/test_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_n$N.log
tail -5 gpurun_out/test_n$N.log
( time timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ) 2> gpurun_out/bench_n$N.time; echo "bench rc=$?"
tail -c 5000 gpurun_out/bench_n$N.json; tail -8 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.time
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 2 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err ) 2> gpurun_out/bench_ref_n$N.time; echo "ref rc=$?"
tail -c 1500 gpurun_out/bench_ref_n$N.json; cat gpurun_out/bench_ref_n$N.time
