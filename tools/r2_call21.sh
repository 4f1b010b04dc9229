#!/bin/bash
# round-2 GPU call 21 (one GPU): S tuples per join work item (table rebuilds of hot partitions) on the Zipf workload and C1
set -u
mkdir -p gpurun_out
for w in c5_zipf c1; do timeout 300 bash tools/sweep_variants.sh $w; done > gpurun_out/r2c21_sweep.log 2>&1; cat gpurun_out/r2c21_sweep.log
