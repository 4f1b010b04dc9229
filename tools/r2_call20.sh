#!/bin/bash
# round-2 GPU call 20 (one GPU): warp-aggregated ranks in the scatter passes of a skewed probe side -- parity, A/B on C5, C1, C3
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py -m gpu -x -q > $o/r2c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r2c20_pytest.log
for w in c5_zipf c1 c3; do timeout 300 bash tools/sweep_variants.sh $w; done > $o/r2c20_sweep.log 2>&1; cat $o/r2c20_sweep.log
HWBRJ_TRACE=1 timeout 200 python tools/prof_c1.py c5_zipf 3 2>&1 | tail -16 | grep -E "scatter|hist|zipf|join"
