#!/bin/bash
# sweep K2 residency knobs on C1 (development helper)
for ctas in 2 3 4 5 6 8; do for rp in 2 4; do
  echo "ctas=$ctas ranges=$rp: $(HWBRJ_PROBE_CTAS=$ctas HWBRJ_RANGE_PASSES=$rp python tools/prof_c1.py c1 3 | tail -1)"
done; done
