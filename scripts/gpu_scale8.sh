#!/bin/bash
# Eight B200s of one box: split mode against one GPU and the oracle, the default bench line at N = 8 (cfg3 strong scaling
# + the split_cfg4 sub-object); cfg5 at full size with HALMA_SCALE8_CFG5=1.
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_scale8.sh'
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
    scripts/split_check.py --giant 2000000 2>gpurun_out/split_check8.err | tee gpurun_out/split_check_n8.txt
tail -2 gpurun_out/split_check8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 \
    bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/bench_n8.err | grep '^{' | tail -1 | tee gpurun_out/bench_default_n8.json | cut -c1-300
tail -3 gpurun_out/bench_n8.err
if [ -n "$HALMA_SCALE8_CFG5" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 \
    bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 3 --no-one-sided --e2e-steps 1 2>gpurun_out/bench_cfg5_n8.err | grep '^{' | tail -1 | tee gpurun_out/bench_cfg5_n8.json | cut -c1-300
tail -3 gpurun_out/bench_cfg5_n8.err
fi
