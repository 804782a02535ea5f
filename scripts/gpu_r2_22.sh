#!/bin/bash
# round 2: the small kernels of the step after batching their per-read metadata (compact, purge flag / compact, fill_rem): launch list of one step + bench line
mkdir -p gpurun_out
COMMON="--no-e2e --no-cpu-baseline --extras= --no-autotune --no-ascii-leg --no-edges --multi-k 0"
timeout 600 python bench.py --steps 5 $COMMON > gpurun_out/bench22.json 2> gpurun_out/bench22.err; echo "rc=$?"; tail -c 300 gpurun_out/bench22.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches22.csv python bench.py --steps 2 --warmup 3 $COMMON > gpurun_out/ncu_launches22.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json, csv
d = json.loads(open("gpurun_out/bench22.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), d["kernels_ms"], d["check"]["checksum_total"], d["check"]["n_solid_total"])
lines = [l for l in open("gpurun_out/launches22.csv") if l.startswith('"')]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
    seq.append((row["Kernel Name"][:50], v))
idx = [i for i, (n, v) in enumerate(seq) if "sketch_packed" in n and v > 10]
for n, v in seq[idx[-2]:idx[-1]]: print(f"{v:9.4f} ms  {n}")
PY
