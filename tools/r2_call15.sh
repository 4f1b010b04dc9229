#!/bin/bash
# round-2 GPU call 15 (two GPUs): multi-GPU parity of the final build (one group and four pipelined groups), bench lines
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_driver.py -m gpu -x -q > $o/r2c15_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -4 $o/r2c15_pytest_dist.log
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps $2 --warmup 3 --e2e-steps $3 --workload $4; }
timeout 300 bash -c "$(declare -f run2); run2 29542 10 3 c1" > $o/r2c15_bench_c1_2gpu.json 2> $o/r2c15_bench_c1_2gpu.err; echo "bench rc=$?"; cut -c1-250 $o/r2c15_bench_c1_2gpu.json; tail -2 $o/r2c15_bench_c1_2gpu.err
timeout 300 bash -c "$(declare -f run2); run2 29543 5 1 c5_zipf" > $o/r2c15_bench_c5_zipf_2gpu.json 2> $o/r2c15_bench_c5_zipf_2gpu.err; echo "zipf rc=$?"; cut -c1-250 $o/r2c15_bench_c5_zipf_2gpu.json
timeout 200 python bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $o/r2c15_ref_2gpu.json 2>&1; echo "ref arm (N=2 flag, no torchrun) rc=$?"; cut -c1-200 $o/r2c15_ref_2gpu.json
