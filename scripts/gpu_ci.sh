#!/bin/bash
# What a GPU trip normally runs (under gpurun, one GPU): GPU parity tests, smoke, both bench arms, cfg4,
# the ncu launch list and full captures of the potential kernel (reduced and full cfg2 size):
#   gpurun --timeout 900 -- 'bash scripts/gpu_ci.sh'
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_reference.json | cut -c1-200
timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/bench.json | cut -c1-300
timeout 300 python bench.py --workload cfg4 --steps 2 --warmup 3 2>/dev/null | tee gpurun_out/bench_cfg4_n1_reuse.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_reuse.csv \
    python scripts/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_potential_fast -c 3 \
    -o gpurun_out/prof_reuse python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1
NCU_STARS=200000 NCU_GAS=500000 timeout 400 ncu --set full --clock-control none -k regex:k_potential_fast -c 1 \
    -o gpurun_out/prof_reuse_fullsize python scripts/ncu_target.py > gpurun_out/ncu_fullsize.log 2>&1
ls -la gpurun_out
