#!/bin/bash
# A/B of builds of the library: tools/ab_libs.sh libA.so libB.so ...   (SHAPES="b c h w;..." overrides the shapes)
IFS=';' read -ra SH <<< "${SHAPES:-256 256 56 56;256 512 28 28;256 1024 14 14;256 64 56 56;256 128 28 28;256 256 14 14}"
for lib in "$@"; do
  for shape in "${SH[@]}"; do
    echo -n "$lib: "; DCTP_LIB=$PWD/dct_pruning_b200/$lib timeout 120 python tools/prof_one.py $shape ${PATHSEL:-auto} 10 2>&1 | tail -1
  done
done
