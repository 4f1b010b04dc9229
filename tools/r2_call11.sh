#!/bin/bash
# round-2 GPU call 11 (one GPU): parity after the scatter / K2 changes, traces, range-pass sweep of the BLOCKED probe,
# the reference's measurement grids through the C driver (CSV), and an `ncu --set full` capture of one C1 join.
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $o/r2c11_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r2c11_pytest_gpu.log
HWBRJ_TRACE=1 timeout 200 python tools/prof_c1.py c1 3 > $o/r2c11_trace_c1.log 2>&1; tail -14 $o/r2c11_trace_c1.log
HWBRJ_TRACE=1 timeout 200 python tools/prof_c1.py c5_zipf 3 > $o/r2c11_trace_c5_zipf.log 2>&1; tail -14 $o/r2c11_trace_c5_zipf.log | grep -E "K2|zipf"
HWBRJ_PROBE_ADAPTIVE=0 timeout 200 python tools/prof_c1.py c5_zipf 3 2>&1 | tail -1
HWBRJ_PROBE_ADAPTIVE=0 timeout 200 python tools/prof_c1.py c1 3 2>&1 | tail -1
timeout 300 python tools/sweep_ranges.py c1_blocked 1,2,4 > $o/r2c11_ranges_c1_blocked.log 2>&1; cat $o/r2c11_ranges_c1_blocked.log
timeout 200 python tools/sweep_ranges.py c1 2,4 > $o/r2c11_ranges_c1.log 2>&1; cat $o/r2c11_ranges_c1.log
for grid in smoke never_single_pass best_bloom_filter_type basic_vs_blocked; do
  timeout 600 python tools/run_sweep.py --grid $grid --out $o/r2c11_sweep_$grid.csv > $o/r2c11_sweep_$grid.log 2>&1; echo "sweep $grid rc=$? rows=$(wc -l < $o/r2c11_sweep_$grid.csv)"
done
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 16 --launch-count 14 -f -o $o/r2c11_c1_full \
  python tools/prof_c1.py c1 2 > $o/r2c11_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $o/r2c11_ncu.log; ls -la $o/r2c11_c1_full.ncu-rep
