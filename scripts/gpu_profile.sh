#!/bin/bash
# ncu launch list of the bench command + one full capture of the sketch kernel and of the insert kernel.
mkdir -p gpurun_out
nvidia-smi -i 0 --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits > gpurun_out/smi_query.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -s 3 -c 1 -f -o gpurun_out/prof_sketch python bench.py --reads 200000 --steps 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_sketch.log 2>&1; echo "sketch capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 3 -c 1 -f -o gpurun_out/prof_insert python bench.py --reads 200000 --steps 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_insert.log 2>&1; echo "insert capture rc=$?"
ls -la gpurun_out
