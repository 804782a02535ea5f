#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --reads 300000 --steps 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 1800 gpurun_out/bench_n$N.json; grep -v "OMP_NUM\|\*\*\*" gpurun_out/bench_n$N.err | tail -5
