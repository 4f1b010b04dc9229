#!/bin/bash
# development helper: time every tuning variant (build/variants/lib_*.so, compiled with -DHWBRJ_* overrides) on a workload.
# A "_c<N>" in the file name sets HWBRJ_PROBE_CTAS=N (probe CTAs per SM), "_o<NN>" sets HWBRJ_PROBE_CARVEOUT=NN (percent).
w=${1:-c1}
for lib in build/variants/lib_*.so; do
  ctas=$(basename $lib | sed -n 's/.*_c\([0-9][0-9]*\)\(_.*\)\?\.so/\1/p')
  carve=$(basename $lib | sed -n 's/.*_o\([0-9]*\)\(_.*\)\?\.so/\1/p')
  echo "$(basename $lib): $(HWBRJ_PROBE_CARVEOUT=${carve:--1} HWBRJ_PROBE_CTAS=${ctas:-0} HWBRJ_LIB=$PWD/$lib python tools/prof_c1.py $w 3 | tail -1 | sed 's/matches=\([0-9]*\) filtered=\(-\?[0-9]*\) /[\1 \2] /')"
done
