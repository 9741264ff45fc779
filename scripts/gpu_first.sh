#!/bin/bash
# First GPU trip: device info, microbench, smoke, pytest -m gpu, quick throughput probe.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/microbench.txt
from pyhalma_b200 import _lib
print(_lib.device_info(0)); print(_lib.microbench(0))
PY
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
timeout 600 python scripts/probe_perf.py 2>&1 | tee gpurun_out/probe_perf.txt
