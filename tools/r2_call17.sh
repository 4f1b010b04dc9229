#!/bin/bash
# round-2 GPU call 17 (one GPU): compile-time variants of the scatter kernel (tile size, threads, ring depth) on C1 and C3
set -u
mkdir -p gpurun_out
for w in c1 c3; do timeout 300 bash tools/sweep_variants.sh $w; done > gpurun_out/r2c17_sweep_scatter.log 2>&1; cat gpurun_out/r2c17_sweep_scatter.log
