#!/bin/bash
# round 2: launch list + full ncu captures of the kernels of one step / multi-k / edges (1 GPU)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --extras '' > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench_quick rc=$?"
tail -c 3000 gpurun_out/bench_quick.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --extras '' --multi-k 6 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sketch_packed_kernel|insert_kernel|pack_ascii_kernel" -s 6 -c 3 -f -o gpurun_out/prof_step python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras '' --multi-k 0 --no-edges > gpurun_out/ncu_step.log 2>&1; echo "step capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"next_k_kernel|edge_insert_kernel|edge_values_kernel|edge_emit_kernel|table_emit_kernel|table_stats_kernel" -c 10 -f -o gpurun_out/prof_table python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras '' --multi-k 5 --no-ascii-leg > gpurun_out/ncu_table.log 2>&1; echo "table capture rc=$?"
ls -la gpurun_out
