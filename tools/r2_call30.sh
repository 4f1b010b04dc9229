#!/bin/bash
# round-2 GPU call 30 (one GPU): per-launch trace of one C1 join of the final build
mkdir -p gpurun_out
HWBRJ_TRACE=1 timeout 40 python tools/prof_c1.py c1 3 > gpurun_out/r2c30_trace_c1.log 2>&1; tail -19 gpurun_out/r2c30_trace_c1.log
