#!/bin/bash
# round-2 GPU call 22 (one GPU): last check of the committed build -- full GPU suite, smoke, default bench line
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $o/r2c22_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r2c22_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/r2c22_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $o/r2c22_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > $o/r2c22_bench_c1.json 2> $o/r2c22_bench_c1.err; echo "bench rc=$?"; cut -c1-300 $o/r2c22_bench_c1.json
