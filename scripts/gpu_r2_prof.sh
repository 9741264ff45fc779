#!/bin/bash
# Round 2: ncu capture of the first (full) potential pass of the cfg3 catalogue (stand-alone kernel, enqueue driver),
# and the e2e overlap sweep.
mkdir -p gpurun_out
HALMA_DRIVER=enqueue timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_potential_fast -c 1 \
    -o gpurun_out/prof_cfg3_pass0 python scripts/cfg3_parts.py --one 0/1 --steps 1 > gpurun_out/ncu_cfg3_pass0.log 2>&1
tail -3 gpurun_out/ncu_cfg3_pass0.log
HALMA_DRIVER=enqueue timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_potential_fast -c 1 \
    -o gpurun_out/prof_cfg3_pass0_r0of8 python scripts/cfg3_parts.py --one 0/8 --steps 1 > gpurun_out/ncu_cfg3_pass0_r0of8.log 2>&1
for s in 1 3 6 8 12; do
  timeout 300 python bench.py --steps 3 --warmup 3 --reps 4 --no-sub --no-cpu --no-one-sided --e2e-streams $s 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('streams', $s, 'e2e ms/run %.2f vs resident %.3f value %.0f'%(d['e2e']['ms_per_run'], d['e2e']['vs_resident'], d['value']))"
done 2>&1 | tee gpurun_out/e2e_sweep.txt
