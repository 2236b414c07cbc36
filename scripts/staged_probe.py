#!/usr/bin/env python3
"""Per-launch overhead of the decode kernel: a staged step (one launch per phase, 142 launches) vs the fused step."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))
import torch
from qwen_megakernel import model_tts
from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to
torch.cuda.set_device(0)
w = weights_to(synthetic_tts_weights(max_seq_len=256), "cuda")
x = synthetic_inputs(99, 16).cuda()
for mode in (0, 1):
    dec = model_tts.TTSDecoder(weights=w, verbose=False, max_seq_len=256, mode=mode)
    for i in range(4):
        dec.step_with_embed(x[i])
    dec._hidden.copy_(x[0])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    a.record()
    for _ in range(n):
        dec._launch(-1, dec._hidden.data_ptr())
    b.record(); torch.cuda.synchronize()
    print(f"mode={mode}: {a.elapsed_time(b)/n*1e3:.1f} us per step", "(142 launches)" if mode else "(1 launch)")
