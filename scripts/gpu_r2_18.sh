#!/bin/bash
# round 2: block form vs warp form of the insert pass at 400 k / 625 k reads (cfg4's shard at N = 8)
mkdir -p gpurun_out
for r in 400000 625000; do for v in 0 1; do
  MDBG_PASS_VARIANT=$v timeout 300 python bench.py --reads $r --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --no-edges --multi-k 0 --steps 5 > gpurun_out/bench18_${r}_v$v.json 2> gpurun_out/bench18_${r}_v$v.err; echo "rc=$?"
  python - $r $v <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench18_{sys.argv[1]}_v{sys.argv[2]}.json").read().strip().splitlines()[-1])
occ = d["check"]["kminmer_occurrences_total"]
print("reads", sys.argv[1], "variant", sys.argv[2], "insert ms", round(d["kernels_ms"]["insert"], 4), "G windows/s", round(occ / d["kernels_ms"]["insert"] / 1e6, 1), d["table_phase_ms_profiled_step_rank0"])
PY
done; done
