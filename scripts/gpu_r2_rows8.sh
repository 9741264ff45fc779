#!/bin/bash
# A/B of the symmetric tickets' row unit (4 vs 8 members per lane) + the GPU suite with the new default.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_rows8.txt
for r in 4 8; do
  echo "== HALMA_SYM_ROWS=$r"
  HALMA_SYM_ROWS=$r timeout 200 python scripts/ncu_single_halo.py 2>&1 | tail -2
  HALMA_SYM_ROWS=$r timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-one-sided --e2e-steps 2 2>gpurun_out/bench_rows$r.err | tee gpurun_out/bench_rows$r.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('cfg3 catalogue_wall_ms %.3f frac %.3f e2e vs resident %.3f'%(d['catalogue_wall_ms'], d['roofline']['frac'], d['e2e']['vs_resident']))
for k in ('cfg2','single_halo_1e6','f2py_level'):
    print(k, json.dumps(d.get(k))[:600])"
done
