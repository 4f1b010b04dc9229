#!/bin/bash
# round-2 GPU call 4 (two GPUs): per-kernel trace of the 2-GPU join (eager), effect of longer level-1 runs over NVLink
set -u
mkdir -p gpurun_out
o=gpurun_out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 1; }
HWBRJ_TRACE=1 timeout 300 bash -c "$(declare -f run2); run2 29541" > $o/r2c4_trace_2gpu.json 2> $o/r2c4_trace_2gpu.err; echo "trace rc=$?"
grep "rank 0" $o/r2c4_trace_2gpu.err | tail -40
HWBRJ_TRACE=1 HWBRJ_RADIX_BITS=12 HWBRJ_L1_BITS=5 timeout 300 bash -c "$(declare -f run2); run2 29542" > $o/r2c4_trace_2gpu_l1b5.json 2> $o/r2c4_trace_2gpu_l1b5.err; echo "trace l1=5 rc=$?"
grep "rank 0" $o/r2c4_trace_2gpu_l1b5.err | tail -24
HWBRJ_TRACE=1 timeout 300 python tools/prof_c1.py c1 3 2>&1 | tail -20
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
