#!/bin/bash
# round 2, GPU call 2: parity tests, full N=1 bench line (with the cfg3/cfg4/cfg5 extras), reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.csv 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test.log
tail -5 gpurun_out/test.log
timeout 300 scripts/k1_probe 1000000 15000 > gpurun_out/k1_probe.json 2> gpurun_out/k1_probe.err; echo "k1_probe rc=$?"; cat gpurun_out/k1_probe.json
( time timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_full.time; echo "bench_full rc=$?"
tail -c 6000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.time
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; echo "bench_ref rc=$?"
tail -c 2500 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.time
ls -la gpurun_out
