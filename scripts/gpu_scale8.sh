#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # n workload port
  if [ "$1" = "1" ]; then timeout 900 python bench.py --workload $2 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_$2_n$1.json
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --workload $2 --steps 3 --warmup 3 2>/dev/null | grep '^{' | tail -1 > gpurun_out/scale_$2_n$1.json; fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_$2_n$1.json")); print("$2 n=$1 value %.0f ms/step %.1f e2e %.0f frac %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["frac"]))
except Exception as e: print("$2 n=$1 FAILED", e)
PY
}
run 8 cfg3 29541; run 4 cfg3 29542; run 2 cfg3 29543; run 1 cfg3 0
run 8 cfg4 29544; run 1 cfg4 0
run 8 cfg2 29545; run 4 cfg2 29546
