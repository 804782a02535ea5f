#!/bin/bash
# round 2: GPU tests of the table paths + quick N = 1 bench (no e2e / cpu baseline) after the abandon-flag fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_new_paths.py tests/test_gpu_zz_round2.py -m gpu -q -x --tb=short > gpurun_out/test5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test5.log; tail -4 gpurun_out/test5.log
timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench5_n1q.json 2> gpurun_out/bench5_n1q.err; echo "bench rc=$?"; tail -3 gpurun_out/bench5_n1q.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench5_n1q.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "kernels", d["kernels_ms"], "phases", d["table_phase_ms_profiled_step_rank0"])
print("multi_k", d["multi_k"]["ms_total"], d["multi_k"]["ms_per_k"])
print("edges", d["edges"]["ms"], "unitigs", {k: v for k, v in d["unitigs"].items() if k != "timer"})
for k, v in d["extras"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
PY
