#!/bin/bash
# What a GPU trip normally runs (under gpurun): GPU parity tests, smoke, both bench arms,
# and the ncu launch list + full capture of the potential kernel.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_ci.sh'
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_reference.json | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/bench.json | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python scripts/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_potential_fast -c 2 \
    -o gpurun_out/prof python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
