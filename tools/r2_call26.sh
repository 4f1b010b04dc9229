#!/bin/bash
# round-2 GPU call 26 (one GPU): K2 with two launch shapes picked by k_probe_sample -- parity, then every workload with the
# sample on and off (off = the shape for dense survivors only)
set -u
mkdir -p gpurun_out
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py -m gpu -x -q > $o/r2c26_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r2c26_pytest.log
for w in c1 c0 c5_zipf c1_blocked_k1; do
  timeout 100 python tools/prof_c1.py $w 3 2>&1 | tail -1
  HWBRJ_PROBE_ADAPTIVE=0 timeout 100 python tools/prof_c1.py $w 3 2>&1 | tail -1 | sed 's/^/   sample off: /'
done 2>&1 | tee $o/r2c26_shapes.log
