#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 ./pyhalma_b200/csrc/pipebench 20000 2>&1 | tee gpurun_out/pipebench_r01.jsonl
for v in 0 1 2 3 4 5 6 7 8 9 10; do
  HALMA_FAST_VARIANT=$v timeout 120 python scripts/probe_variants.py 2>&1 | tail -1 | tee -a gpurun_out/variants.txt
done
timeout 600 python -m pytest tests/test_gpu_unbind.py -q -m gpu -k "idempotence" 2>&1 | tail -3
