#!/bin/bash
# A/B of two builds of the library on the three big ResNet-50 shapes: tools/ab_libs.sh libA.so libB.so [path]
for lib in "$@"; do
  for shape in "256 256 56 56" "256 512 28 28" "256 1024 14 14" "256 64 56 56" "256 128 28 28" "256 256 14 14"; do
    echo -n "$lib: "; DCTP_LIB=$PWD/dct_pruning_b200/$lib timeout 120 python tools/prof_one.py $shape stack 10 2>&1 | tail -1
  done
done
