#!/bin/bash
for h in 1 0; do for w in c1 c0; do
  echo "hashpart=$h $w: $(HWBRJ_HASH_PARTITION=$h python tools/prof_c1.py $w 4 | tail -1)"
done; done
