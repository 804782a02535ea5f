#!/bin/bash
# round 2, GPU call 1: packed sketch kernel on hardware -- probe, parity tests, ncu of the new kernels, one bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
nproc > gpurun_out/nproc.txt
timeout 300 scripts/k1_probe 1000000 15000 > gpurun_out/k1_probe.json 2> gpurun_out/k1_probe.err; echo "k1_probe rc=$?"; cat gpurun_out/k1_probe.json
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test.log
tail -5 gpurun_out/test.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sketch_packed|pack_ascii" -s 2 -c 2 -f -o gpurun_out/prof_packed scripts/k1_probe 1000000 15000 > gpurun_out/ncu_packed.log 2>&1; echo "packed capture rc=$?"
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full rc=$?"
tail -c 2500 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
ls -la gpurun_out
