#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_split.py -q -m gpu 2>&1 | tail -5
timeout 600 $TR --master-port 29531 bench.py --gpus 2 --workload cfg3 --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/bench_cfg3_n2_np.json | cut -c1-150
export HALMA_CFG4_N=500000
timeout 600 $TR --master-port 29532 bench.py --gpus 2 --workload cfg4 --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/bench_cfg4_n2_np.json | cut -c1-150
timeout 600 $TR --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/bench_cfg2_n2_np.json | cut -c1-150
