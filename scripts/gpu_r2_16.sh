#!/bin/bash
# round 2: insert pass, block form vs warp form (both with the single-copy hash and side-by-side loads); baseline 1.78 ms / multi-k 31.6 ms
mkdir -p gpurun_out
for v in 0 1; do
  MDBG_PASS_VARIANT=$v timeout 600 python bench.py --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --no-edges --steps 5 > gpurun_out/bench16_v$v.json 2> gpurun_out/bench16_v$v.err; echo "rc=$?"; tail -c 300 gpurun_out/bench16_v$v.err
  python - $v <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench16_v{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("variant", sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "insert ms", round(d["kernels_ms"]["insert"], 4), "multi_k", d["multi_k"]["ms_total"], d["multi_k"]["ms_per_k"][:3], d["table_phase_ms_profiled_step_rank0"], "checksum", d["check"]["checksum_total"], d["check"]["n_solid_total"])
PY
done
