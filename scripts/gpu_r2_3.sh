#!/bin/bash
# round 2, third session: whole GPU suite on the tree with the unitig kernels, then the default bench (N = 1, every leg)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi3.csv 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test3.log; tail -6 gpurun_out/test3.log
( time timeout 900 python bench.py > gpurun_out/bench3_n1.json 2> gpurun_out/bench3_n1.err ) 2> gpurun_out/bench3_n1.time; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench3_n1.err; cat gpurun_out/bench3_n1.time
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench3_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("packed_host_input"))
print("roofline", {k: d["roofline"][k] for k in ("frac", "traffic", "int_issue")})
print("multi_k", d["multi_k"]["ms_total"], "edges", d["edges"], "unitigs", d["unitigs"])
print("cpu", d["cpu_baseline"])
PY
