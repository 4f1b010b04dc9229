#!/bin/bash
# One-GPU round-end validation: GPU parity tests, smoke, both bench arms, and the ncu launch list of the bench command.
# Outputs go to gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/val_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/val_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/val_smoke.log
timeout 300 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/val_bench_ref.json 2> gpurun_out/val_bench_ref.err; echo "ref rc=$?"
timeout 120 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/val_bench_short.json 2>/dev/null; rc=$?; echo "short bench rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/val_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/val_ncu_bench.log 2>&1; echo "ncu rc=$?"
fi
cut -c1-400 gpurun_out/val_bench.json
