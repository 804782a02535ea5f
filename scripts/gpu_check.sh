#!/bin/bash
# One gpurun call: parity tests, small + full bench, ncu launch list and one full capture
# of the sketch kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test.log
tail -3 gpurun_out/test.log
timeout 300 python bench.py --reads 100000 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench_small rc=$?"
tail -c 1500 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full rc=$?"
tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?"
tail -c 1200 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
