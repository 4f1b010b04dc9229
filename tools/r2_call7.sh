#!/bin/bash
# round-2 GPU call 7 (two GPUs): pipelined groups of owned bins on two streams -- parity, then bench with 1 / 2 / 4 groups
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py tests/test_gpu_knobs.py -m gpu -x -q > $o/r2c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $o/r2c7_pytest.log
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 1; }
for parts in 1 2 4; do
  HWBRJ_DIST_PARTS=$parts timeout 300 bash -c "$(declare -f run2); run2 2954$parts" > $o/r2c7_bench_c1_2gpu_parts$parts.json 2> $o/r2c7_bench_c1_2gpu_parts$parts.err; echo "parts=$parts rc=$?"; cut -c1-200 $o/r2c7_bench_c1_2gpu_parts$parts.json
done
HWBRJ_TRACE=1 timeout 300 python tools/prof_c1.py c1 3 2>&1 | tail -15
