#!/bin/bash
# GPU check of the reuse options (external-sum cache, incremental passes):
#   gpurun --timeout 900 -- 'bash scripts/gpu_reuse_check.sh'
# 1. their own tests (tests/test_gpu_reuse.py, including the forced-on battery of test_gpu_unbind.py)
# 2. the default bench line and the catalogue with the options on
# 3. the whole GPU suite with the options on by default
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_reuse.py -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_reuse.txt
timeout 300 python bench.py --reuse 1 --steps 3 --warmup 3 2>gpurun_out/bench_cfg2_reuse.err | tee gpurun_out/bench_cfg2_reuse.json | cut -c1-400
timeout 300 python bench.py --reuse 1 --workload cfg3 --steps 3 --warmup 3 2>gpurun_out/bench_cfg3_reuse.err | tee gpurun_out/bench_cfg3_reuse.json | cut -c1-400
HALMA_CACHE_EXT=1 HALMA_INCREMENTAL=1 timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_reuse.py 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_reuse_default.txt
ls -la gpurun_out
