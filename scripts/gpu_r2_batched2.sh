#!/bin/bash
export QMK_TIMEOUT_CYCLES=600000000
timeout 900 python -m pytest tests/test_gpu_batched.py -m gpu -x -q > gpurun_out/r2_batched_tests.log 2>&1; echo "pytest rc=$?"
grep -E "^E|passed|failed|batched B" gpurun_out/r2_batched_tests.log | head -20
timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
from qwen_megakernel.model_tts import BatchedFrameLoop
from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to
w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=2048), "cuda")
for B in (16, 64):
    loop = BatchedFrameLoop(w, B, max_seq_len=512)
    pre = synthetic_inputs(1, 8 * B).cuda().view(8, B, 1024)
    extra = synthetic_inputs(2, B).cuda()
    loop.start(pre)
    for _ in range(3): loop.frame(extra)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    a.record()
    for _ in range(n): loop.frame(extra)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"B={B}: {ms:.3f} ms per frame for all streams = {B * 1000 / ms:.0f} codec frames/s per GPU", flush=True)
    del loop
PY
