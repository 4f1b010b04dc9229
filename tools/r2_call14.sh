#!/bin/bash
# round-2 GPU call 14 (one GPU): round-end validation of the final build -- full GPU suite, smoke, both bench arms,
# ncu launch list of the bench command
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $o/r2c14_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r2c14_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/r2c14_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $o/r2c14_smoke.log
timeout 400 python bench.py > $o/r2c14_bench_c1.json 2> $o/r2c14_bench_c1.err; echo "bench rc=$?"; cut -c1-300 $o/r2c14_bench_c1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $o/r2c14_bench_c1_reference.json 2> $o/r2c14_bench_c1_reference.err; echo "ref rc=$?"; cut -c1-300 $o/r2c14_bench_c1_reference.json
for w in c3 c5_zipf c0 c1_blocked; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $o/r2c14_bench_$w.json 2> $o/r2c14_bench_$w.err; echo "bench $w rc=$?"; cut -c1-200 $o/r2c14_bench_$w.json
done
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $o/r2c14_bench_short.json 2>/dev/null; rc=$?; echo "short bench rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2c14_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $o/r2c14_ncu_bench.log 2>&1; echo "ncu rc=$?"
fi
