#!/bin/bash
for rep in 1 2; do
  for t in 0 1; do
    QMK_AUTOTUNE=$t QMK_CALIBRATE_VERBOSE=$t timeout 200 python scripts/probe2.py --configs 2 --positions 100,300 2> gpurun_out/tune_$t.err | grep "^cfg" | sed "s/^cfg= *2/autotune=$t/"
  done
done
grep "autotune group" gpurun_out/tune_1.err | head -8
QMK_AUTOTUNE=1 timeout 200 python scripts/probe2.py --configs 2 --positions 100 --trace 2>/dev/null | grep -E "per-group|per layer"
