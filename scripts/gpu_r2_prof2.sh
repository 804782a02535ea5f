#!/bin/bash
# round 2: launch list of two steps + multi-k (k=4..6) + edges, and full captures of insert / next-k / edge kernels
mkdir -p gpurun_out
COMMON="--no-e2e --no-cpu-baseline --extras= --no-autotune --no-ascii-leg"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 $COMMON --multi-k 6 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^insert_kernel|insert_kernel<" -s 3 -c 1 -f -o gpurun_out/prof_insert python bench.py --steps 1 --warmup 3 $COMMON --multi-k 0 --no-edges > gpurun_out/ncu_insert.log 2>&1; echo "insert capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"next_k_kernel" -s 2 -c 2 -f -o gpurun_out/prof_nextk python bench.py --steps 1 --warmup 3 $COMMON --multi-k 21 --no-edges > gpurun_out/ncu_nextk.log 2>&1; echo "next-k capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"edge_" -s 3 -c 3 -f -o gpurun_out/prof_edges python bench.py --steps 1 --warmup 3 $COMMON --multi-k 0 > gpurun_out/ncu_edges.log 2>&1; echo "edges capture rc=$?"
ls -la gpurun_out | tail -12
