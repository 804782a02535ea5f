#!/bin/bash
# round 2: the N = 2 GPU tests (real NCCL) + the N = 2 bench (parity leg, headline, cfg4 / cfg5) on the tree with warp-form passes and free table capacities
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short > gpurun_out/test20_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test20_n2.log; tail -3 gpurun_out/test20_n2.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench20_n2.json 2> gpurun_out/bench20_n2.err ) 2> gpurun_out/bench20_n2.time; echo "bench rc=$?"
grep -v "^\*\|^Setting\|^$\|^W" gpurun_out/bench20_n2.err | tail -5; cat gpurun_out/bench20_n2.time
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench20_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], (d["e2e"].get("packed_host_input") or {}).get("value"))
print("parity", d["check"]["multi_gpu_parity"])
print("multi_k", d["multi_k"]["ms_total"], "edges", d["edges"]["ms"])
for k, v in d["extras"].items():
    print(k, v.get("value"), v.get("ms_per_step"), v.get("merge_at_first_and_last_k_only"))
PY
