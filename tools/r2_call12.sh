#!/bin/bash
# round-2 GPU call 12 (one GPU): parity after the K2 two-loop / skew-sample change and the join's parked matches, A/B of
# the join variants on three workloads, skew switch on/off, ncu --set full of the K2 launches of the final build
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py -m gpu -x -q > $o/r2c12_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r2c12_pytest.log
for w in c1 c3 c5_zipf c0; do timeout 300 bash tools/sweep_variants.sh $w; done > $o/r2c12_sweep_join.log 2>&1; cat $o/r2c12_sweep_join.log
HWBRJ_TRACE=1 timeout 200 python tools/prof_c1.py c1 3 > $o/r2c12_trace_c1.log 2>&1; tail -15 $o/r2c12_trace_c1.log
HWBRJ_TRACE=1 timeout 200 python tools/prof_c1.py c5_zipf 3 > $o/r2c12_trace_c5_zipf.log 2>&1; tail -15 $o/r2c12_trace_c5_zipf.log | grep -E "K2|skew|zipf"
HWBRJ_PROBE_ADAPTIVE=0 timeout 200 python tools/prof_c1.py c5_zipf 3 2>&1 | tail -1
HWBRJ_PROBE_ADAPTIVE=0 timeout 200 python tools/prof_c1.py c1 3 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_probe_compact --launch-skip 2 --launch-count 2 -f -o $o/r2c12_k2_full \
  python tools/prof_c1.py c1 2 > $o/r2c12_ncu.log 2>&1; echo "ncu rc=$?"; tail -1 $o/r2c12_ncu.log
