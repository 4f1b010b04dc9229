#!/bin/bash
# round-2 GPU call 13 (one GPU): scatter write-out through warp shuffles -- parity, then A/B against the shared-memory
# write-out on four workloads, trace of C1
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py -m gpu -x -q > $o/r2c13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r2c13_pytest.log
for w in c1 c3 c5_zipf c0; do timeout 300 bash tools/sweep_variants.sh $w; done > $o/r2c13_sweep_scatter.log 2>&1; cat $o/r2c13_sweep_scatter.log
HWBRJ_TRACE=1 timeout 200 python tools/prof_c1.py c1 3 > $o/r2c13_trace_c1.log 2>&1; tail -15 $o/r2c13_trace_c1.log
HWBRJ_TRACE=1 timeout 200 python tools/prof_c1.py c5_zipf 3 2>&1 | tail -15 | grep -E "scatter|zipf"
