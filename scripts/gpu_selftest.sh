#!/bin/bash
# Seconds-long GPU check of the newest paths without importing torch; results in gpurun_out/selftest_*.json.
# Build first (here, no GPU needed): see the nvcc lines below.  Usage on the box: bash scripts/gpu_selftest.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/scripts/fake_nccl_cuda:$LD_LIBRARY_PATH
timeout 12 scripts/e2e_probe 262144 15000 4 > gpurun_out/selftest_e2e_default.json 2>&1
timeout 20 python scripts/gpu_selftest.py variants > gpurun_out/selftest_variants.json 2>&1
timeout 20 python scripts/gpu_selftest.py pieces > gpurun_out/selftest_pieces.json 2>&1
timeout 20 python scripts/gpu_selftest.py ranks 2 > gpurun_out/selftest_ranks2.json 2>&1
MDBG_PIECE_PIPELINE=0 timeout 10 scripts/e2e_probe 262144 15000 4 > gpurun_out/selftest_e2e_nopipeline.json 2>&1
timeout 20 python scripts/gpu_selftest.py ranks 3 > gpurun_out/selftest_ranks3.json 2>&1
MDBG_PACK_ISA=avx2 timeout 10 scripts/e2e_probe 262144 15000 4 > gpurun_out/selftest_e2e_avx2.json 2>&1
tail -n 3 gpurun_out/selftest_*.json
