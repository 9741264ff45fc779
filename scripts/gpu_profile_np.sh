#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/bench_cfg3_np.json | cut -c1-150
HALMA_CFG4_N=500000 timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/bench_cfg4_np.json | cut -c1-150
timeout 600 python bench.py --workload cfg1 --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/bench_cfg1_np.json | cut -c1-150
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv \
    python scripts/ncu_target.py > gpurun_out/ncu_launches_b.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_potential_fast -c 2 \
    -o gpurun_out/prof_np_r01 python scripts/ncu_target.py > gpurun_out/ncu_full_b.log 2>&1
tail -3 gpurun_out/ncu_full_b.log
