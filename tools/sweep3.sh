#!/bin/bash
for o in 1 0; do for w in c1 c0 c1_blocked; do
  echo "overlap=$o $w: $(HWBRJ_OVERLAP=$o python tools/prof_c1.py $w 4 | tail -1)"
done; done
