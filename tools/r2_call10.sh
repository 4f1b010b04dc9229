#!/bin/bash
# round-2 GPU call 10 (eight GPUs): parity at 4/8 GPUs, per-kernel trace, bench lines of every workload at 8 GPUs, C1 at 4
set -u
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi topo -m > $o/r2c10_topo.txt 2>&1
lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" > $o/r2c10_numa.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_driver.py -m gpu -x -q > $o/r2c10_pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -4 $o/r2c10_pytest_dist.log
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps $3 --warmup 3 --e2e-steps $4 --workload $5; }
HWBRJ_TRACE=1 timeout 200 bash -c "$(declare -f runN); runN 8 29541 3 1 c1" > $o/r2c10_trace_8gpu.json 2> $o/r2c10_trace_8gpu.err; echo "trace rc=$?"
grep "rank 0" $o/r2c10_trace_8gpu.err | tail -39
for w in c1 c5_zipf c0 c3 c1_blocked; do
  timeout 200 bash -c "$(declare -f runN); runN 8 29551 20 3 $w" > $o/r2c10_bench_${w}_8gpu.json 2> $o/r2c10_bench_${w}_8gpu.err; echo "bench8 $w rc=$?"; cut -c1-230 $o/r2c10_bench_${w}_8gpu.json
done
timeout 200 bash -c "$(declare -f runN); runN 4 29547 20 3 c1" > $o/r2c10_bench_c1_4gpu.json 2> $o/r2c10_bench_c1_4gpu.err; echo "bench4 rc=$?"; cut -c1-230 $o/r2c10_bench_c1_4gpu.json
for parts in 1 2; do
  HWBRJ_DIST_PARTS=$parts timeout 200 bash -c "$(declare -f runN); runN 8 2956$parts 20 1 c1" > $o/r2c10_bench_c1_8gpu_parts$parts.json 2> $o/r2c10_bench_c1_8gpu_parts$parts.err; echo "parts=$parts rc=$?"; cut -c1-230 $o/r2c10_bench_c1_8gpu_parts$parts.json
done
