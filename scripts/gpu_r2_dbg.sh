#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4
timeout 120 python scripts/passes.py cfg1 2>&1 | tail -12
timeout 200 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu --no-one-sided 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('cfg1 ms/step %.3f e2e ms %.3f'%(d['ms_per_step'], d['e2e']['ms_per_run']), d['roofline']['phase_ms_per_run'])"
timeout 200 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('cfg2 ms/step %.3f value %.0f frac %.3f one-sided %.3f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['one_sided']['roofline_frac']))"
