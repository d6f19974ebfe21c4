#!/bin/bash
# A/B of the score kernels on ResNet-50's layer shapes (batch 256), one line per (shape, path):  tools/perf_shapes.sh [paths...]
PATHS=${@:-"stack tmem"}
for shape in "256 256 56 56" "256 64 56 56" "256 512 28 28" "256 128 28 28" "256 1024 14 14" "256 256 14 14" "256 2048 7 7" "256 512 7 7"; do
  set -- $shape
  for p in $PATHS; do
    if [ "$3" -le 8 ]; then q=$([ "$p" = stack ] && echo kron || echo umma); else q=$p; fi
    timeout 120 python tools/prof_one.py $shape $q 10 2>&1 | tail -1
  done
done
