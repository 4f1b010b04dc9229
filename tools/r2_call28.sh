#!/bin/bash
# round-2 GPU call 28 (two GPUs): multi-GPU parity after the K2 shape change
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "two_gpu or 2" > gpurun_out/r2c28_pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c28_pytest_dist.log
