#!/bin/bash
# After a change to the kernels or the scheduler (one GPU, ~1.3 GPU-minutes): the GPU suite, the warp-busy fractions of
# the potential phases on the whole cfg3 catalogue and on one GPU's eighth of it, and the per-pass phase times of
# cfg1 / cfg2 / cfg4.   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_check.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_check.txt
bash scripts/gpu_r2_busy.sh 2>&1 | grep -E "pass [0-3]:|per_pass" | cut -c1-420 | tee gpurun_out/busy_check.txt
for w in cfg1 cfg2 cfg4; do timeout 200 python scripts/passes.py $w 2>&1 | grep -E "total|pass [01]"; done | tee gpurun_out/passes_check.txt
