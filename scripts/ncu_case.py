#!/usr/bin/env python3
"""Small fixed workload for ncu captures (round 2): 8 prefill steps + step(BOS) [talker launches 0..8], one fused
code-predictor frame [launch 9], one talker step_with_codes [10], then generate_frames(3) [11: the frame-loop kernel, 3 frames
in one launch], talker steps at positions 500 and 2047 [12, 13], and (with --batched) two batched steps at B = 64."""
import os
import sys

os.environ.setdefault("QMK_AUTOTUNE", "0")   # the in-situ autotuner adds ~100 decode launches at model creation: keep the launch indices fixed

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))
import torch
from qwen_megakernel.model_tts import BatchedTTSDecoder, CodePredictorKernel, TTSDecoder
from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to

w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=2048), "cuda")
dec = TTSDecoder(weights=w, verbose=False, max_seq_len=2048)
cp = CodePredictorKernel(w, device="cuda")
x = synthetic_inputs(99, 12).cuda()
for i in range(8):
    dec.step_with_embed(x[i])
tok, hid = dec.step(2149)
codes = cp.predict(hid, tok, w["embed_weight"], do_sample=True, temperature=0.9, top_k=50)
tok, hid = dec.step_with_codes(codes, cp.codec_embeddings, x[8])
out = dec.generate_frames(cp, 3, x[9:12], x[0], do_sample=True, temperature=0.9, top_k=50, eos_token=-1)
for pos in (500, 2047):
    dec._position = pos
    dec.step_with_embed(x[1])
if "--batched" in sys.argv:
    for pers in ("0", "1"):
        os.environ["QMK_BATCHED_PERSISTENT"] = pers
        bd = BatchedTTSDecoder(w, 64, max_seq_len=512)
        t = torch.full((64,), 2149, dtype=torch.int32, device="cuda")
        for _ in range(2):
            t2, _ = bd.step(t)
        torch.cuda.synchronize()
        del bd
torch.cuda.synchronize()
print("ncu_case done", out[2])
