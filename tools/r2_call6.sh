#!/bin/bash
# round-2 GPU call 6 (eight GPUs): in-process multi-GPU parity, per-kernel trace at 8 GPUs, C1 at 8 and 4 GPUs, Zipf at 8
set -u
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi topo -m > $o/r2c6_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $o/r2c6_pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -5 $o/r2c6_pytest_dist.log
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps $3 --warmup 3 --e2e-steps $4 --workload $5; }
HWBRJ_TRACE=1 timeout 300 bash -c "$(declare -f runN); runN 8 29541 3 1 c1" > $o/r2c6_trace_8gpu.json 2> $o/r2c6_trace_8gpu.err; echo "trace rc=$?"
grep "rank 0" $o/r2c6_trace_8gpu.err | tail -21
grep "rank 5" $o/r2c6_trace_8gpu.err | tail -21
timeout 300 bash -c "$(declare -f runN); runN 8 29542 20 3 c1" > $o/r2c6_bench_c1_8gpu.json 2> $o/r2c6_bench_c1_8gpu.err; echo "bench8 rc=$?"; cut -c1-220 $o/r2c6_bench_c1_8gpu.json
timeout 300 bash -c "$(declare -f runN); runN 4 29543 20 3 c1" > $o/r2c6_bench_c1_4gpu.json 2> $o/r2c6_bench_c1_4gpu.err; echo "bench4 rc=$?"; cut -c1-220 $o/r2c6_bench_c1_4gpu.json
timeout 400 bash -c "$(declare -f runN); runN 8 29544 5 1 c5_zipf" > $o/r2c6_bench_c5_zipf_8gpu.json 2> $o/r2c6_bench_c5_zipf_8gpu.err; echo "zipf8 rc=$?"; cut -c1-220 $o/r2c6_bench_c5_zipf_8gpu.json; tail -3 $o/r2c6_bench_c5_zipf_8gpu.err
