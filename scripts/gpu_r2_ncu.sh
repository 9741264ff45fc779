#!/bin/bash
# Round-2 ncu evidence: launch list of the default bench command (reduced repetitions) and full captures of the
# persistent loop kernel on the cfg3 catalogue and on one 1e6-star halo, and of the first (full) potential pass of
# the catalogue as a stand-alone kernel.  The metric pages and the executed-MUFU counts are extracted on the box;
# only the catalogue capture is kept as .ncu-rep (gpurun_out/ travels back up to 64 MiB).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --reps 2 --no-sub --no-cpu --no-one-sided --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
extract() {      # name, analytic evaluations
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  python scripts/ncu_mufu_count.py gpurun_out/$1.ncu-rep $2 > gpurun_out/mufu_$1.json 2>gpurun_out/mufu_$1.err
  head -c 600 gpurun_out/mufu_$1.json; echo
}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unbind_loop --launch-skip 1 -c 1 \
    -o gpurun_out/prof_loop_cfg3 -f python scripts/cfg3_parts.py --one 0/1 --steps 1 > gpurun_out/ncu_loop_cfg3.log 2>&1
tail -1 gpurun_out/ncu_loop_cfg3.log | cut -c1-600
extract prof_loop_cfg3 88186432805
NCU_N=1000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unbind_loop --launch-skip 1 -c 2 \
    -o gpurun_out/prof_loop_1e6 -f python scripts/ncu_single_halo.py > gpurun_out/ncu_loop_1e6.log 2>&1
tail -2 gpurun_out/ncu_loop_1e6.log
extract prof_loop_1e6 500063997952
rm -f gpurun_out/prof_loop_1e6.ncu-rep
HALMA_DRIVER=enqueue timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_potential_fast -c 1 \
    -o gpurun_out/prof_cfg3_pass0 -f python scripts/cfg3_parts.py --one 0/1 --steps 1 > gpurun_out/ncu_cfg3_pass0.log 2>&1
tail -1 gpurun_out/ncu_cfg3_pass0.log | cut -c1-300
extract prof_cfg3_pass0 77805642133
rm -f gpurun_out/prof_cfg3_pass0.ncu-rep
du -sh gpurun_out
