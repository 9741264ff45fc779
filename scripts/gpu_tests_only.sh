#!/bin/bash
# All GPU tests, nothing else:  gpurun --timeout 400 -- 'bash scripts/gpu_tests_only.sh'
mkdir -p gpurun_out
timeout 380 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
