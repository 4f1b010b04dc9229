#!/bin/bash
# times every tuning variant in build/variants on C1 (development helper)
for lib in build/variants/lib_*.so; do
  for c in 2 3 4 6; do
  echo "$(basename $lib) ctas=$c: $(HWBRJ_PROBE_CTAS=$c HWBRJ_LIB=$PWD/$lib python tools/prof_c1.py c1 3 | tail -1 | sed 's/matches=[0-9]* filtered=-\?[0-9]* //')"
  done
done
