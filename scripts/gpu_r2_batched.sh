#!/bin/bash
set -x
export QMK_TIMEOUT_CYCLES=600000000
timeout 900 python -m pytest tests/test_gpu_batched.py -m gpu -x -q > gpurun_out/r2_batched_tests.log 2>&1; echo "pytest rc=$?"
grep -E "^E|passed|failed|batched B" gpurun_out/r2_batched_tests.log | head -20
timeout 300 python - <<'PY' > gpurun_out/r2_batched_time.log 2>&1
import os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
import bench
from qwen_megakernel.synthetic import synthetic_tts_weights, weights_to
w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=2048), "cuda")
for chain in ("0", "1"):
    os.environ["QMK_BATCHED_PERSISTENT"] = "0" if chain == "1" else "1"
    for b in (16, 64):
        r = bench.time_batched(w, torch.device("cuda", 0), b)
        print("chain" if chain == "1" else "persistent", b, round(r["ms_per_step"], 4), "ms/step", round(r["stream_steps_per_s"]), "stream-steps/s", round(r["algorithmic_gbs"]), "GB/s", flush=True)
PY
cat gpurun_out/r2_batched_time.log | tail -8
