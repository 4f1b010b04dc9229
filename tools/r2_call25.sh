#!/bin/bash
# round-2 GPU call 25 (one GPU): new K2 shape (4 keys per lane, 5 CTAs/SM, 256-tuple rings) -- parity, then A/B against
# the old shape on every workload whose probe runs k_probe_compact
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py -m gpu -x -q > $o/r2c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r2c25_pytest.log
for w in c1 c0 c5_zipf c1_blocked_k1 c1_blocked; do timeout 300 bash tools/sweep_variants.sh $w; done > $o/r2c25_sweep.log 2>&1; cat $o/r2c25_sweep.log
