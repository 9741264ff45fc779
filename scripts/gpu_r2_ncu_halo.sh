#!/bin/bash
# ncu capture of the persistent loop kernel on ONE 1e6-star halo (symmetric pass with pairs of row tiles, then the
# one-sided pass); metric page and executed-MUFU count extracted on the box.
mkdir -p gpurun_out
NCU_N=1000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unbind_loop --launch-skip 1 -c 2 \
    -o gpurun_out/prof_loop_1e6 -f python scripts/ncu_single_halo.py > gpurun_out/ncu_loop_1e6.log 2>&1
tail -2 gpurun_out/ncu_loop_1e6.log
ncu -i gpurun_out/prof_loop_1e6.ncu-rep --page raw --csv > gpurun_out/prof_loop_1e6_raw.csv 2>/dev/null
python scripts/ncu_mufu_count.py gpurun_out/prof_loop_1e6.ncu-rep 500063997952 > gpurun_out/mufu_prof_loop_1e6.json 2>gpurun_out/mufu_prof_loop_1e6.err
head -c 500 gpurun_out/mufu_prof_loop_1e6.json
