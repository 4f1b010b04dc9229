#!/bin/bash
# round-2 GPU call 18 (one GPU): K2 variants (warps per CTA, TMA bulk prefetch of S into L2), scatter delta, join unroll
set -u
mkdir -p gpurun_out
for w in c1; do timeout 400 bash tools/sweep_variants.sh $w; done > gpurun_out/r2c18_sweep.log 2>&1; cat gpurun_out/r2c18_sweep.log
