#!/bin/bash
for t in 1 0; do for c in 3 4 5 6; do
  echo "tma=$t ctas=$c c1: $(HWBRJ_PROBE_TMA=$t HWBRJ_PROBE_CTAS=$c python tools/prof_c1.py c1 4 | tail -1)"
done; done
for t in 1 0; do echo "tma=$t c0: $(HWBRJ_PROBE_TMA=$t python tools/prof_c1.py c0 4 | tail -1)"; done
