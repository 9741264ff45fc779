#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unbind.py -x -q -m gpu -k "cfg2_full or cfg3_full or lattice or duplicates" 2>&1 | tail -5
for v in 0 2 5 6; do
  echo "variant $v"; HALMA_FAST_VARIANT=$v timeout 300 python bench.py --workload cfg4 --steps 2 --warmup 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('  value %.0f  kernel %.0f  frac %.3f'%(d['value'], r['achieved'], r['frac']))"
done
