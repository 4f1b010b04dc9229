#!/bin/bash
# round-2 GPU call 29 (two GPUs): C1 bench line on 2 GPUs after the K2 shape change
mkdir -p gpurun_out
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/r2c29_bench_c1_2gpu.json 2> gpurun_out/r2c29_bench_c1_2gpu.err; echo "rc=$?"; cut -c1-220 gpurun_out/r2c29_bench_c1_2gpu.json
