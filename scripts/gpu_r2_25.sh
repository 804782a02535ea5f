#!/bin/bash
# round 2, last call: smoke() + a short bench on the final tree (the table-pass roofline object, allocation counters)
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --steps 5 > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo "rc=$?"; tail -c 300 gpurun_out/bench25.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench25.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "multi_k", d["multi_k"]["ms_total"], "edges", d["edges"]["ms"], "unitigs", d["unitigs"]["ms"])
print({k: v for k, v in d["roofline_table_pass"].items() if k != "note"})
PY
