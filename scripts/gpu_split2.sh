#!/bin/bash
# Split mode with incremental passes on 2 GPUs: gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_split2.sh'
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_split.py -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_split.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 2 --workload cfg4 --steps 2 --warmup 3 2>gpurun_out/bench_cfg4_n2.err | grep '^{' | tail -1 | tee gpurun_out/bench_cfg4_n2_reuse.json | cut -c1-300
tail -5 gpurun_out/bench_cfg4_n2.err
