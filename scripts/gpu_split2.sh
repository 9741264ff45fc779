#!/bin/bash
# Split mode and the N = 2 default bench line: gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_split2.sh'
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_split.py -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_split.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    scripts/split_check.py --giant 2000000 2>gpurun_out/split_check.err | tee gpurun_out/split_check.txt
tail -3 gpurun_out/split_check.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 \
    bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2.err | grep '^{' | tail -1 | tee gpurun_out/bench_default_n2.json | cut -c1-400
tail -5 gpurun_out/bench_n2.err
