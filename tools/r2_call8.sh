#!/bin/bash
# round-2 GPU call 8 (one GPU): state of HEAD after the container restart -- full GPU suite, smoke, per-kernel trace and
# bench lines of every single-GPU workload. Outputs: gpurun_out/r2c8_*.
set -u
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $o/r2c8_gpu.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/r2c8_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $o/r2c8_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $o/r2c8_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $o/r2c8_pytest_gpu.log
HWBRJ_TRACE=1 timeout 300 python tools/prof_c1.py c1 3 > $o/r2c8_trace_c1.log 2>&1; tail -16 $o/r2c8_trace_c1.log
for w in c0 c3 c1_blocked c5_zipf; do
  HWBRJ_TRACE=1 timeout 300 python tools/prof_c1.py $w 3 > $o/r2c8_trace_$w.log 2>&1; tail -1 $o/r2c8_trace_$w.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > $o/r2c8_bench_c1.json 2> $o/r2c8_bench_c1.err; echo "bench c1 rc=$?"; cut -c1-400 $o/r2c8_bench_c1.json
for w in c0 c3 c1_blocked c5_zipf; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $o/r2c8_bench_$w.json 2> $o/r2c8_bench_$w.err
  echo "bench $w rc=$?"; cut -c1-300 $o/r2c8_bench_$w.json
done
