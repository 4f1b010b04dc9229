#!/bin/bash
# times every tuning variant in build/variants on the partition-heavy C3 and on C1 (development helper)
for lib in build/variants/lib_*.so; do
  echo "$(basename $lib) c3: $(HWBRJ_LIB=$PWD/$lib python tools/prof_c1.py c3 3 | tail -1 | sed 's/matches=[0-9]* filtered=-\?[0-9]* //')"
  echo "$(basename $lib) c1: $(HWBRJ_LIB=$PWD/$lib python tools/prof_c1.py c1 3 | tail -1 | sed 's/matches=[0-9]* filtered=-\?[0-9]* //')"
done
