#!/bin/bash
# round 2, final tree: N-rank tests + bench at N = $1 (default 2), and the default N = 1 bench on the same box
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi8_n$N.csv 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_round2.py -m gpu -q -x --tb=short > gpurun_out/test8_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test8_n$N.log
tail -4 gpurun_out/test8_n$N.log
( time timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench8_n$N.json 2> gpurun_out/bench8_n$N.err ) 2> gpurun_out/bench8_n$N.time; echo "bench N=$N rc=$?"
tail -4 gpurun_out/bench8_n$N.err; cat gpurun_out/bench8_n$N.time
if [ "$2" = "with-n1" ]; then
  ( time timeout 900 python bench.py > gpurun_out/bench8_n1.json 2> gpurun_out/bench8_n1.err ) 2> gpurun_out/bench8_n1.time; echo "bench N=1 rc=$?"
  ( time timeout 900 python bench.py --impl reference > gpurun_out/bench8_ref.json 2> gpurun_out/bench8_ref.err ) 2> gpurun_out/bench8_ref.time; echo "ref rc=$?"
fi
python - $N <<'PY'
import json, sys
for f in (f"bench8_n{sys.argv[1]}", "bench8_n1", "bench8_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "missing", e); continue
    print(f, "value", d["value"], "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "packed", ((d.get("e2e") or {}).get("packed_host_input") or {}).get("value"))
    print("   multi_k", (d.get("multi_k") or {}).get("ms_total"), "edges", (d.get("edges") or {}).get("ms"), "unitigs", (d.get("unitigs") or {}).get("ms"), "parity", (d.get("check") or {}).get("multi_gpu_parity"))
    for k, v in (d.get("extras") or {}).items():
        print("   ", k, v.get("value"), v.get("ms_per_step"), (v.get("merge_at_first_and_last_k_only") or {}).get("value"), v.get("error"))
PY
