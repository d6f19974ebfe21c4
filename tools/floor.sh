#!/bin/bash
# per-launch cost of the score kernels on launches too small to be bandwidth bound (back-to-back launches, no events between)
for shape in "12 16 10 10" "12 16 20 20" "12 16 40 40" "12 64 40 40" "256 16 32 32" "256 64 8 8" "256 32 16 16" "12 16 80 80" "128 512 2 2" "256 12 32 32"; do
  set -- $shape
  for p in auto umma; do
    echo -n "$p: "; timeout 120 python tools/prof_one.py $shape $p 200 2>&1 | tail -1
  done
done
