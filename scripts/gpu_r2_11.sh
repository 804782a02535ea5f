#!/bin/bash
# round 2: where does the k = 4 step of the multi-k extra spend 20 ms after the unitig extra?
mkdir -p gpurun_out
for tag in edges noedges; do
  extra=""; [ $tag = noedges ] && extra="--no-edges"
  timeout 600 python bench.py --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --steps 3 $extra > gpurun_out/bench11_$tag.json 2> gpurun_out/bench11_$tag.err; echo "rc=$?"
  python - $tag <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench11_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "value", d["value"], "multi_k", d["multi_k"]["ms_total"], d["multi_k"]["ms_per_k"][:4], "first sweep", d["multi_k"]["ms_per_k_first_sweep"][:4])
PY
done
