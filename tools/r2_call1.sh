#!/bin/bash
# round-2 GPU call 1 (one GPU): validate the opt-in paths of round 1, time the K2 ablation / tuning builds, and take
# bench lines of the non-headline workloads. Outputs: gpurun_out/r2c1_*.
set -u
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $o/r2c1_gpu.txt 2>&1
HWBRJ_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q -x > $o/r2c1_pytest_experimental.log 2>&1
echo "experimental rc=$?"; tail -3 $o/r2c1_pytest_experimental.log
# K2 ablation (results of b/c/d are wrong by construction: timing only)
rm -rf build/variants_k2 && mv build/variants build/variants_k2 && mv build/variants_ablate build/variants
timeout 400 bash tools/sweep_variants.sh c1 > $o/r2c1_sweep_ablate.log 2>&1; cat $o/r2c1_sweep_ablate.log
mv build/variants build/variants_ablate && mv build/variants_k2 build/variants
timeout 900 bash tools/sweep_variants.sh c1 > $o/r2c1_sweep_k2.log 2>&1; cat $o/r2c1_sweep_k2.log
for w in c0 c3 c1_blocked c5_zipf; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $o/r2c1_bench_$w.json 2> $o/r2c1_bench_$w.err
  echo "bench $w rc=$?"; cut -c1-300 $o/r2c1_bench_$w.json
done
