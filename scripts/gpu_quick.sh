#!/bin/bash
# Quick GPU check after a change to the loop: all GPU tests, then the cfg2 and cfg3 bench lines.
#   gpurun --timeout 600 -- 'bash scripts/gpu_quick.sh'
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json | cut -c1-300
