#!/bin/bash
# A/B of kernel variants (variants/libqmk_*.so) with scripts/probe2.py, two alternating rounds
for rep in 1 2; do
  for v in "$@"; do
    QMK_LIB_PATH=$PWD/variants/libqmk_$v.so timeout 200 python scripts/probe2.py --configs 2 --positions ${POSITIONS:-100,300,1000} 2>&1 | grep "^cfg" | sed "s/^cfg= *2/$v/"
  done
done
