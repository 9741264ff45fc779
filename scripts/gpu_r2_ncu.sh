#!/bin/bash
# Round-2 ncu evidence: launch list of the default bench command (reduced repetitions) and full captures of the
# persistent loop kernel on the cfg3 catalogue and on one 1e6-star halo.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --reps 2 --no-sub --no-cpu --no-one-sided --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unbind_loop --launch-skip 1 -c 1 \
    -o gpurun_out/prof_loop_cfg3 python scripts/cfg3_parts.py --one 0/1 --steps 1 > gpurun_out/ncu_loop_cfg3.log 2>&1
tail -2 gpurun_out/ncu_loop_cfg3.log
ncu -i gpurun_out/prof_loop_cfg3.ncu-rep --page raw --csv > gpurun_out/prof_loop_cfg3_raw.csv 2>/dev/null
NCU_N=1000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unbind_loop --launch-skip 1 -c 2 \
    -o gpurun_out/prof_loop_1e6 python scripts/ncu_single_halo.py > gpurun_out/ncu_loop_1e6.log 2>&1
tail -2 gpurun_out/ncu_loop_1e6.log
ncu -i gpurun_out/prof_loop_1e6.ncu-rep --page raw --csv > gpurun_out/prof_loop_1e6_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep | tail -4
