#!/bin/bash
# Second GPU trip: full GPU suite, bench (both arms), ncu launch list + full capture.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_cfg2.json
tail -5 gpurun_out/bench_err.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/bench_err.txt | tee gpurun_out/bench_ref.json
timeout 600 python bench.py --workload cfg1 --steps 5 --warmup 3 2>> gpurun_out/bench_err.txt | tee gpurun_out/bench_cfg1.json
# ncu: launch list of one short bench run (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python scripts/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
# ncu: full capture of the fast potential kernel (3 launches)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_potential_fast -c 2 \
    -o gpurun_out/prof_fast_r01 python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
