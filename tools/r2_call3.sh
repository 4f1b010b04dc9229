#!/bin/bash
# round-2 GPU call 3 (two GPUs): multi-GPU parity (torchrun ranks over IPC, in-process GPUs over peer access), 2-GPU bench
set -u
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi topo -m > $o/r2c3_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_driver.py -m gpu -x -q > $o/r2c3_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -25 $o/r2c3_pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $o/r2c3_bench_c1_2gpu.json 2> $o/r2c3_bench_c1_2gpu.err; echo "bench2 rc=$?"; cut -c1-1500 $o/r2c3_bench_c1_2gpu.json; tail -5 $o/r2c3_bench_c1_2gpu.err
for i in 1 2; do timeout 300 python tools/prof_c1.py c1 4 | tail -1; done
