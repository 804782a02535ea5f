#!/bin/bash
# parity tests + device-resident bench only (no e2e / cpu baseline), for kernel iteration
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test.log
tail -4 gpurun_out/test.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1])
    print("value", round(j["value"],1), "Gbp/s  ms/step", round(j["ms_per_step"],2), "kernels", j["kernels_ms"], "roofline frac", round(j["roofline"]["frac"],4), "launches", j["gpu_launches"], "clocks", j["clocks"])
except Exception as e:
    print("parse fail", e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
