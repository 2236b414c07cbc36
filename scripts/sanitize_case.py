#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): a 2-layer talker, one fused code-predictor
frame (greedy + sampled), the fused embedding-sum step, a long-context step with split attention and a 1-layer batched step."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))

import torch  # noqa: E402

from qwen_megakernel.model_tts import BatchedTTSDecoder, CodePredictorKernel, TTSDecoder  # noqa: E402
from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to  # noqa: E402

torch.cuda.set_device(0)
w = weights_to(synthetic_tts_weights(num_layers=2, max_seq_len=256), "cuda")
x = synthetic_inputs(5, 8).cuda()
dec = TTSDecoder(weights=w, verbose=False, max_seq_len=256)
tok, hid = dec.step(2149)
tok, hid = dec.step_with_embed(x[0])
cp = CodePredictorKernel(w, device="cuda")
codes = cp.predict(hid, tok, w["embed_weight"], do_sample=False)
codes2 = cp.predict(hid, tok, w["embed_weight"], do_sample=True, temperature=0.9, top_k=50)
tok, hid = dec.step_with_codes(codes, cp.codec_embeddings, x[1])
dec._position = 130                      # split attention (S = 4): rows 0..129 are zeros from the allocation
tok, hid = dec.step_with_embed(x[2])
bd = BatchedTTSDecoder(w, 16, max_seq_len=32, num_layers=1)
t, h = bd.step(torch.full((16,), 2149, dtype=torch.int32, device="cuda"))
t2, h2 = bd.step(t)
torch.cuda.synchronize()
print("sanitize case ok", tok, codes.tolist()[:4], codes2.tolist()[:4], t2[:4].tolist())
