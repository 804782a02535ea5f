#!/bin/bash
# round 2: whole GPU suite + default bench on the tree with the unitig graph edges
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test10.log; tail -5 gpurun_out/test10.log
( time timeout 900 python bench.py > gpurun_out/bench10_n1.json 2> gpurun_out/bench10_n1.err ) 2> gpurun_out/bench10_n1.time; echo "bench rc=$?"; tail -c 600 gpurun_out/bench10_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench10_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], (d["e2e"].get("packed_host_input") or {}).get("value"))
print("multi_k", d["multi_k"]["ms_total"], "edges", d["edges"]["ms"])
print("unitigs", {k: v for k, v in d["unitigs"].items() if k != "timer"})
PY
