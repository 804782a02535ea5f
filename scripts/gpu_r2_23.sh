#!/bin/bash
# round 2, final tree: the default bench at N = 8 (parity leg over real NCCL, headline, e2e, multi-k, cfg4 / cfg5 strong scaling)
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/bench23_n8.json 2> gpurun_out/bench23_n8.err ) 2> gpurun_out/bench23_n8.time; echo "bench rc=$?"
grep -v "^\*\|^Setting\|^$\|^W" gpurun_out/bench23_n8.err | tail -5; cat gpurun_out/bench23_n8.time
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench23_n8.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], (d["e2e"].get("packed_host_input") or {}).get("value"))
print("parity", {k: v for k, v in d["check"]["multi_gpu_parity"].items() if k != "what"})
print("multi_k", d["multi_k"]["ms_total"], "edges", d["edges"]["ms"], d["table_phase_ms_profiled_step_rank0"])
for k, v in d["extras"].items():
    print(k, v.get("value"), v.get("ms_per_step"), (v.get("merge_at_first_and_last_k_only") or {}).get("value"), (v.get("merge_at_first_and_last_k_only") or {}).get("ms_per_step"), v.get("checksum_total_last_k"))
PY
