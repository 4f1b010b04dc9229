#!/bin/bash
# round-2 GPU call 5 (two GPUs): pull-based level 2 -- parity, trace, bench
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -m gpu -x -q > $o/r2c5_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $o/r2c5_pytest.log
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps $2 --warmup 3 --e2e-steps 1; }
HWBRJ_TRACE=1 timeout 300 bash -c "$(declare -f run2); run2 29541 3" > $o/r2c5_trace_2gpu.json 2> $o/r2c5_trace_2gpu.err; echo "trace rc=$?"
grep "rank 0" $o/r2c5_trace_2gpu.err | tail -21
timeout 300 bash -c "$(declare -f run2); run2 29542 10" > $o/r2c5_bench_c1_2gpu.json 2> $o/r2c5_bench_c1_2gpu.err; echo "bench rc=$?"; cut -c1-200 $o/r2c5_bench_c1_2gpu.json
HWBRJ_TRACE=1 timeout 300 python tools/prof_c1.py c1 3 2>&1 | tail -15
