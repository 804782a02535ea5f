#!/bin/bash
# round 2, final tree: whole GPU suite, smoke(), the default bench and the reference arm on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test24.log; tail -3 gpurun_out/test24.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke24.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke24.log
( time timeout 900 python bench.py > gpurun_out/bench24_n1.json 2> gpurun_out/bench24_n1.err ) 2> gpurun_out/bench24_n1.time; echo "bench rc=$?"; tail -c 400 gpurun_out/bench24_n1.err
( time timeout 900 python bench.py --impl reference > gpurun_out/bench24_ref.json 2> gpurun_out/bench24_ref.err ) 2> gpurun_out/bench24_ref.time; echo "ref rc=$?"; tail -c 300 gpurun_out/bench24_ref.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench24_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], (d["e2e"].get("packed_host_input") or {}).get("value"), "allocs", d["device_allocations_in_timed_region"])
print("multi_k", d["multi_k"]["ms_total"], d["multi_k"]["allocations_rank0"], "edges", d["edges"]["ms"], "unitigs", d["unitigs"]["ms"], d["unitigs"]["cpu_reference_on_sample"]["identical_records"], d["unitigs"]["cpu_reference_on_sample"]["identical_edge_lists"])
for k, v in d["extras"].items(): print(k, v.get("value"), v.get("ms_per_step"))
print("cpu_baseline", d["cpu_baseline"])
r = json.loads(open("gpurun_out/bench24_ref.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["cpu_baseline"]["cores"], r.get("reference_multi_k", {}).get("seconds_total"))
PY
