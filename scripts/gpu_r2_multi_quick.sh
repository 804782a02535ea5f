#!/bin/bash
# multi-GPU bench only (parity leg + headline + extras), N = $1
N=${1:-2}
mkdir -p gpurun_out
( time timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-e2e ${@:2} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ) 2> gpurun_out/bench_n$N.time; echo "bench rc=$?"
tail -c 4000 gpurun_out/bench_n$N.json; grep -v "^\*\|^Setting\|^$" gpurun_out/bench_n$N.err | tail -8; cat gpurun_out/bench_n$N.time
