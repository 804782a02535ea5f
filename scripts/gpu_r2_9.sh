#!/bin/bash
# round 2: e2e leg vs reads per host-buffer call; then a from-scratch build on the GPU box + smoke
mkdir -p gpurun_out
COMMON="--no-cpu-baseline --extras= --no-autotune --no-ascii-leg --multi-k 0 --no-edges --steps 3 --warmup 3"
for B in 262144 524288 1048576; do
  timeout 400 python bench.py --e2e-batch $B $COMMON > gpurun_out/e2e_b$B.json 2> gpurun_out/e2e_b$B.err
  python - $B <<'PY'
import json, sys
b = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/e2e_b{b}.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print(b, "reads/call: e2e", e["value"], "packed input", (e.get("packed_host_input") or {}).get("value"), "pieces", e["last_host_batch"]["n_pieces"], "pack GB/s", e["last_host_batch"]["pack_gb_per_s"], "checks", e["table_checks"])
except Exception as ex:
    print(b, "failed", ex, open(f"gpurun_out/e2e_b{b}.err").read()[-800:])
PY
done
( cd metamdbg_b200/csrc && make -s clean ) ; rm -f metamdbg_b200/libmdbg_b200.so
( time python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/fresh_build_smoke.log 2>&1; echo "fresh build + smoke rc=$?"; tail -3 gpurun_out/fresh_build_smoke.log
