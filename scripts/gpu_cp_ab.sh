#!/bin/bash
# A/B of library variants on the code-predictor frame (greedy / sampled), alternating, two rounds: scripts/gpu_cp_ab.sh base smatch ...
for rep in 1 2; do
  for v in "$@"; do
    QMK_LIB_PATH=$PWD/variants/libqmk_$v.so timeout 120 python scripts/cp_time.py 2>&1 | grep -E "greedy|sampled|codes" | cut -c1-150 | sed "s/^/$v: /"
  done
done
