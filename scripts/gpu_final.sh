#!/bin/bash
# tests + full bench + reference arm + ncu (launch list, sketch + insert captures)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test.log
tail -3 gpurun_out/test.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?"
python - <<'PY'
import json
for f in ("bench_full","bench_ref"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(j["value"],3), "e2e", j["e2e"]["value"], "cpu", j.get("cpu_baseline"), "stages", j.get("reference_stages"))
    except Exception as e:
        print(f, "parse fail", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -s 3 -c 1 -f -o gpurun_out/prof_sketch python bench.py --reads 200000 --steps 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_sketch.log 2>&1; echo "sketch capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -s 3 -c 1 -f -o gpurun_out/prof_sketch_full python bench.py --steps 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_sketch_full.log 2>&1; echo "sketch full-size capture rc=$?"
