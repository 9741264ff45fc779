#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_corr.txt
for w in cfg2 cfg3 cfg4; do timeout 200 python scripts/passes.py $w 2>&1 | tail -14; done | tee gpurun_out/passes_corr.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 2 2>gpurun_out/bench_corr.err | tee gpurun_out/bench_corr.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('cfg3 catalogue_wall_ms %.3f frac %.3f e2e vs resident %.3f'%(d['catalogue_wall_ms'], d['roofline']['frac'], d['e2e']['vs_resident']))
for k in ('cfg2','single_halo_1e6'):
    print(k, json.dumps(d.get(k))[:900])"
