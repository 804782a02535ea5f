#!/bin/bash
# round 2: after the abandon-flag fix -- insert rate vs table size again, launch list, full captures of the table / unitig kernels
mkdir -p gpurun_out
COMMON="--no-e2e --no-cpu-baseline --extras= --no-autotune --no-ascii-leg"
for R in 12500 50000 200000; do
  timeout 300 python bench.py --reads $R $COMMON --multi-k 0 --no-edges --steps 5 --warmup 3 > gpurun_out/ins2_$R.json 2> gpurun_out/ins2_$R.err
  python - $R <<'PY'
import json, sys
r = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ins2_{r}.json").read().strip().splitlines()[-1])
    occ = d["check"]["kminmer_occurrences_total"]
    print(r, "reads: insert ms", d["kernels_ms"]["insert"], "windows", occ, "Gwin/s", occ / d["kernels_ms"]["insert"] / 1e6, "step", d["ms_per_step"])
except Exception as e:
    print(r, "failed", e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches6.csv python bench.py --steps 2 --warmup 3 $COMMON --multi-k 7 > gpurun_out/ncu_launches6.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^insert_kernel|insert_kernel<|next_k_stream_kernel" -s 3 -c 3 -f -o gpurun_out/prof6_table python bench.py --steps 1 --warmup 3 $COMMON --multi-k 7 --no-edges > gpurun_out/ncu6_table.log 2>&1; echo "table capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"unitig_|edge_" -c 40 -f -o gpurun_out/prof6_graph python bench.py --steps 1 --warmup 3 $COMMON --multi-k 0 > gpurun_out/ncu6_graph.log 2>&1; echo "graph capture rc=$?"
ls -la gpurun_out | tail -6
