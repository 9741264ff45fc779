#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_unbind.py -q -x -k "multi_stream or concurrent or ragged" 2>&1 | tail -3
for s in 1 3 4 6 8; do
  timeout 300 python bench.py --steps 3 --warmup 3 --reps 4 --no-sub --no-cpu --no-one-sided --e2e-streams $s 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('streams', $s, 'e2e ms/run %.2f vs resident %.3f value %.0f'%(d['e2e']['ms_per_run'], d['e2e']['vs_resident'], d['value']))"
done 2>&1 | tee gpurun_out/e2e_sweep.txt
timeout 600 python scripts/parity_full.py gpu --out gpurun_out/parity_r02.json > gpurun_out/parity_full.md 2>&1; tail -3 gpurun_out/parity_full.md
