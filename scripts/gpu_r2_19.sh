#!/bin/bash
# round 2: table capacity (power of two vs multiple of 1024 at load factor 0.6) x pass form (block / warp) at cfg2 size
mkdir -p gpurun_out
for p in 1 0; do for v in 0 1; do
  MDBG_TABLE_POW2=$p MDBG_PASS_VARIANT=$v timeout 300 python bench.py --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --no-edges --steps 5 > gpurun_out/bench19_p${p}_v$v.json 2> gpurun_out/bench19_p${p}_v$v.err; echo "rc=$?"; tail -c 300 gpurun_out/bench19_p${p}_v$v.err
  python - $p $v <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench19_p{sys.argv[1]}_v{sys.argv[2]}.json").read().strip().splitlines()[-1])
occ = d["check"]["kminmer_occurrences_total"]
print("pow2", sys.argv[1], "variant", sys.argv[2], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "insert ms", round(d["kernels_ms"]["insert"], 4), "G windows/s", round(occ / d["kernels_ms"]["insert"] / 1e6, 1), "multi_k", d["multi_k"]["ms_total"], d["table_phase_ms_profiled_step_rank0"], d["check"]["checksum_total"])
PY
done; done
