#!/bin/bash
for d in 1 0; do for ctas in 3 4; do
  echo "defer=$d ctas=$ctas: $(HWBRJ_DEFER=$d HWBRJ_PROBE_CTAS=$ctas python tools/prof_c1.py c1 3 | tail -1)"
done; done
echo "defer=1 auto: $(python tools/prof_c1.py c1 3 | tail -1)"
echo "defer=1 ranges4: $(HWBRJ_RANGE_PASSES=4 python tools/prof_c1.py c1 3 | tail -1)"
echo "c1_blocked: $(python tools/prof_c1.py c1_blocked 3 | tail -1)"
