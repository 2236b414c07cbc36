#!/bin/bash
# rebuild libqmk_b200.so and print the decode kernels' resource usage
cd "$(dirname "$0")/.." && python - <<'PY' 2>&1 | grep -E -A3 "decode_kernel(_traced)?EN|rror" | grep -E "Compiling|spill|registers|rror" | head -20
import sys
sys.path.insert(0,'qwen-megakernel-tts_b200')
from qwen_megakernel import build_tts
try:
    build_tts.build(force=True, verbose=True)
except Exception as e:
    print("error", str(e)[-3000:])
PY
