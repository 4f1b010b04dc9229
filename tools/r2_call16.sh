#!/bin/bash
# round-2 GPU call 16 (eight GPUs): final build -- parity at 2/4/8 GPUs, C1 and Zipf at 8, C1 at 4 (two groups = default, and one)
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_driver.py -m gpu -x -q > $o/r2c16_pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -2 $o/r2c16_pytest_dist.log
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps $3 --warmup 3 --e2e-steps $4 --workload $5; }
timeout 120 bash -c "$(declare -f runN); runN 8 29551 20 3 c1" > $o/r2c16_bench_c1_8gpu.json 2> $o/r2c16_bench_c1_8gpu.err; echo "bench8 c1 rc=$?"; cut -c1-230 $o/r2c16_bench_c1_8gpu.json
timeout 120 bash -c "$(declare -f runN); runN 8 29552 10 1 c5_zipf" > $o/r2c16_bench_c5_zipf_8gpu.json 2> $o/r2c16_bench_c5_zipf_8gpu.err; echo "bench8 zipf rc=$?"; cut -c1-230 $o/r2c16_bench_c5_zipf_8gpu.json
timeout 120 bash -c "$(declare -f runN); runN 4 29553 20 1 c1" > $o/r2c16_bench_c1_4gpu.json 2> $o/r2c16_bench_c1_4gpu.err; echo "bench4 rc=$?"; cut -c1-230 $o/r2c16_bench_c1_4gpu.json
HWBRJ_DIST_PARTS=1 timeout 120 bash -c "$(declare -f runN); runN 4 29554 20 1 c1" > $o/r2c16_bench_c1_4gpu_one_group.json 2> $o/r2c16_bench_c1_4gpu_one_group.err; echo "bench4 one group rc=$?"; cut -c1-230 $o/r2c16_bench_c1_4gpu_one_group.json
