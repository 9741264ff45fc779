#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_2gpu.txt
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 2>gpurun_out/err_cfg3_1.txt | tee gpurun_out/bench_cfg3_n1.json
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --workload cfg3 --steps 3 --warmup 3 2>gpurun_out/err_cfg3_2.txt | tee gpurun_out/bench_cfg3_n2.json
export HALMA_CFG4_N=500000
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 2>gpurun_out/err_cfg4_1.txt | tee gpurun_out/bench_cfg4_n1.json
timeout 600 $TR --master-port 29522 bench.py --gpus 2 --workload cfg4 --steps 3 --warmup 3 2>gpurun_out/err_cfg4_2.txt | tee gpurun_out/bench_cfg4_n2.json
timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/err_cfg2_2.txt | tee gpurun_out/bench_cfg2_n2.json
tail -3 gpurun_out/err_*.txt
