#!/bin/bash
# Reduced 8-GPU scaling check with the symmetric kernel: gpurun --gpus 8 -- 'bash scripts/gpu_scale8_sym.sh'
mkdir -p gpurun_out
run() { # n workload port
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --workload $2 --steps 2 --warmup 3 2>/dev/null | grep '^{' | tail -1 > gpurun_out/scale_sym_$2_n$1.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_sym_$2_n$1.json")); o=d["one_sided"]
    print("$2 n=$1 value %.0f ms/step %.1f e2e %.0f eval-frac %.3f inter/eval %.2f | one-sided (%s) value %.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["frac"],d["roofline"]["interactions_per_evaluation"],o["scope"],o["value"]))
except Exception as e: print("$2 n=$1 FAILED", e)
PY
}
run 8 cfg3 29541
run 8 cfg4 29544
