#!/bin/bash
# Round 2: GPU check of the persistent loop kernel (fused driver): smoke, GPU tests, cfg3 partitions, cfg1, default bench line.
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.txt
if ! grep -q "smoke ok" gpurun_out/smoke.txt; then echo "SMOKE FAILED"; exit 1; fi
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
timeout 300 python scripts/cfg3_parts.py --worlds 1,8 --steps 5 2>&1 | tee gpurun_out/cfg3_parts_fused.jsonl | cut -c1-900
timeout 200 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu 2>gpurun_out/cfg1.err | tee gpurun_out/bench_cfg1_fused.json | cut -c1-400
timeout 300 python scripts/bench_f2py_call.py 2>&1 | tee gpurun_out/bench_f2py_call.jsonl | cut -c1-500
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/bench.err | tee gpurun_out/bench_default.json | cut -c1-600
tail -5 gpurun_out/bench.err
