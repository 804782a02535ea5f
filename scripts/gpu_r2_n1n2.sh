#!/bin/bash
# N=1 quick bench (no e2e / cpu baseline) then N=2 (if 2 GPUs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test.log; tail -4 gpurun_out/test.log
timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_n1q.json 2> gpurun_out/bench_n1q.err; echo "n1 rc=$?"; tail -3 gpurun_out/bench_n1q.err
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then bash scripts/gpu_r2_multi_quick.sh $NG > gpurun_out/multi_quick.log 2>&1; tail -3 gpurun_out/multi_quick.log; fi
