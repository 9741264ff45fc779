#!/bin/bash
# Round-2 final single-GPU evidence: smoke, the default bench line (both arms), launch list + full ncu captures.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 600 python bench.py 2>gpurun_out/bench_default_n1.err | tail -1 > gpurun_out/bench_default_n1.json; cut -c1-300 gpurun_out/bench_default_n1.json
timeout 600 python bench.py --impl reference 2>gpurun_out/bench_reference.err | tail -1 > gpurun_out/bench_reference.json; cut -c1-300 gpurun_out/bench_reference.json
bash scripts/gpu_r2_ncu.sh
