#!/bin/bash
# One-GPU check after a kernel change: GPU parity tests, smoke, default bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/val_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/val_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/val_bench.json')); print(d['value'], d['ms_per_step'], d['phases_ms'], d['clocks'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
