#!/bin/bash
# round 2 (final kernels): launch list of two steps + multi-k (k = 4..6) + edges + unitigs; full captures of the warp-form table passes
mkdir -p gpurun_out
COMMON="--no-e2e --no-cpu-baseline --extras= --no-autotune --no-ascii-leg"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches21.csv python bench.py --steps 2 --warmup 3 $COMMON --multi-k 6 > gpurun_out/ncu_launches21.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"insert_warp_kernel|next_k" -s 3 -c 22 -f -o gpurun_out/prof21_passes python bench.py --steps 1 --warmup 3 $COMMON --multi-k 21 --no-edges > gpurun_out/ncu_passes21.log 2>&1; echo "passes capture rc=$?"
ls -la gpurun_out | grep 21
