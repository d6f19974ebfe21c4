#!/bin/bash
# A/B of the stacked-basis kernel's role layouts (DCTP_STACK_CFG) on ResNet-50 shapes: tools/cfg_sweep.sh 4 7 ...
for cfg in ${@:-4 7}; do
  for shape in "256 256 56 56" "256 512 28 28" "256 1024 14 14" "256 64 56 56"; do
    echo -n "cfg $cfg: "; DCTP_STACK_CFG=$cfg timeout 120 python tools/prof_one.py $shape stack 10 2>&1 | tail -1
  done
done
