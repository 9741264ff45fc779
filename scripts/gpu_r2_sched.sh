#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_sched.txt
bash scripts/gpu_r2_busy.sh 2>&1 | grep -E "pass [0-3]:|per_pass" | cut -c1-420 | tee gpurun_out/busy_sched.txt
for w in cfg1 cfg2 cfg4; do timeout 200 python scripts/passes.py $w 2>&1 | grep -E "total|pass [01]"; done | tee gpurun_out/passes_sched.txt
