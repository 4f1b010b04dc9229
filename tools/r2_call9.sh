#!/bin/bash
# round-2 GPU call 9 (two GPUs): multi-GPU parity on the unified pipeline, per-kernel trace and bench lines at 2 GPUs
set -u
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi topo -m > $o/r2c9_topo.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done > $o/r2c9_numa.txt 2>&1
lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" >> $o/r2c9_numa.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_driver.py -m gpu -x -q > $o/r2c9_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -6 $o/r2c9_pytest_dist.log
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps $2 --warmup 3 --e2e-steps $3 --workload $4; }
HWBRJ_TRACE=1 timeout 300 bash -c "$(declare -f run2); run2 29541 3 1 c1" > $o/r2c9_trace_2gpu.json 2> $o/r2c9_trace_2gpu.err; echo "trace rc=$?"
grep "rank 0" $o/r2c9_trace_2gpu.err | tail -28
timeout 300 bash -c "$(declare -f run2); run2 29542 10 3 c1" > $o/r2c9_bench_c1_2gpu.json 2> $o/r2c9_bench_c1_2gpu.err; echo "bench rc=$?"; cut -c1-250 $o/r2c9_bench_c1_2gpu.json; tail -3 $o/r2c9_bench_c1_2gpu.err
timeout 300 bash -c "$(declare -f run2); run2 29543 5 1 c5_zipf" > $o/r2c9_bench_c5_zipf_2gpu.json 2> $o/r2c9_bench_c5_zipf_2gpu.err; echo "zipf rc=$?"; cut -c1-250 $o/r2c9_bench_c5_zipf_2gpu.json; tail -3 $o/r2c9_bench_c5_zipf_2gpu.err
# per-rank sizes of C1 at 8 GPUs, on 2 GPUs (latency-bound regime): r=32M s=256M m=2^28
HWBRJ_TRACE=1 timeout 300 bash -c "$(declare -f run2); run2 29544 3 1 c1_quarter" > $o/r2c9_trace_quarter_2gpu.json 2> $o/r2c9_trace_quarter_2gpu.err; echo "quarter rc=$?"
grep "rank 0" $o/r2c9_trace_quarter_2gpu.err | tail -28; cut -c1-250 $o/r2c9_trace_quarter_2gpu.json
