#!/bin/bash
# Round-2 baseline of the per-pass fixed cost: cfg3 partitions emulated on one GPU, cfg1 step, launch lists.
mkdir -p gpurun_out
timeout 300 python scripts/cfg3_parts.py --worlds 1,8 --steps 5 2>&1 | tee gpurun_out/cfg3_parts_base.jsonl | cut -c1-400
timeout 200 python scripts/cfg3_parts.py --worlds 1,8 --steps 5 --graph 1 2>&1 | tee gpurun_out/cfg3_parts_base_graph.jsonl | cut -c1-400
timeout 200 python bench.py --workload cfg1 --steps 20 --warmup 5 --cpu-targets 1000 2>gpurun_out/cfg1.err | tee gpurun_out/bench_cfg1_base.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg3_r0of8_base.csv \
    python scripts/cfg3_parts.py --one 0/8 --steps 1 > gpurun_out/ncu_cfg3_launches.log 2>&1
tail -3 gpurun_out/ncu_cfg3_launches.log
