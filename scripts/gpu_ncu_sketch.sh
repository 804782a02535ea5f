#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -s 3 -c 1 -f -o gpurun_out/prof_sketch python bench.py --reads 200000 --steps 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_sketch.log 2>&1; echo "sketch capture rc=$?"
