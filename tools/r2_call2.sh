#!/bin/bash
# round-2 GPU call 2 (one GPU): first hardware run of the unified pipeline (single-GPU case), micro-benchmark of gather paths
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 120 build/gather_micro 26 29 > $o/r2c2_gather_micro.log 2>&1; echo "gather rc=$?"; cat $o/r2c2_gather_micro.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/r2c2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $o/r2c2_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $o/r2c2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $o/r2c2_pytest_gpu.log
for w in c1 c0 c3 c1_blocked; do
  timeout 300 python tools/prof_c1.py $w 4 > $o/r2c2_prof_$w.log 2>&1; echo "prof $w rc=$?"; tail -1 $o/r2c2_prof_$w.log
done
HWBRJ_PROBE_STAGED=1 timeout 300 python tools/prof_c1.py c1_blocked 4 > $o/r2c2_prof_c1_blocked_staged.log 2>&1; tail -1 $o/r2c2_prof_c1_blocked_staged.log
HWBRJ_HASH_PARTITION=0 timeout 300 python tools/prof_c1.py c1_blocked 4 > $o/r2c2_prof_c1_blocked_radix.log 2>&1; tail -1 $o/r2c2_prof_c1_blocked_radix.log
