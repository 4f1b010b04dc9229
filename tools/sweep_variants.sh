#!/bin/bash
# development helper: time every tuning variant (build/variants/lib_*.so, compiled with -DHWBRJ_* overrides) on a workload
w=${1:-c1}
for lib in build/variants/lib_*.so; do
  echo "$(basename $lib): $(HWBRJ_LIB=$PWD/$lib python tools/prof_c1.py $w 3 | tail -1 | sed 's/matches=[0-9]* filtered=-\?[0-9]* //')"
done
