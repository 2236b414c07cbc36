#!/bin/bash
# build a kernel variant next to the product library: scripts/build_variant.sh NAME "-DFLAG ..."  -> gpurun_out/.. no: variants/libqmk_NAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
H=$(python -c "
import sys; sys.path.insert(0,'qwen-megakernel-tts_b200')
from qwen_megakernel import build_tts; print(build_tts.source_hash())")
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr --extended-lambda -Xcompiler -fPIC -shared \
  $2 -DQMK_SRC_HASH="\"$H\"" -Iinclude -Iqwen-megakernel-tts_b200/csrc -o variants/libqmk_$1.so \
  qwen-megakernel-tts_b200/csrc/qmk_engine.cu qwen-megakernel-tts_b200/csrc/qmk_batched.cu -Xptxas -v 2>&1 | grep -A2 "qmk218qmk2_decode_kernelE" | grep -E "spill|registers" | tr '\n' ' '
echo " -> variants/libqmk_$1.so"
