"""Batched talker step at depth: us per step with every stream at position p (KV rows of the untouched cache are zeros)."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
from qwen_megakernel.model_tts import BatchedTTSDecoder
from qwen_megakernel.synthetic import synthetic_tts_weights, weights_to
w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=2048), "cuda")
for B in (16, 64):
    dec = BatchedTTSDecoder(w, B, max_seq_len=2048)
    tok = torch.arange(B, dtype=torch.int32, device="cuda") + 5
    out = []
    for p in (0, 100, 500, 1000, 2000):
        for rep in range(2):
            dec.positions.fill_(p); dec._steps = p
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): dec.step(tok)
            b.record(); torch.cuda.synchronize()
        kv_gb = B * 28 * 8 * (p + 5) * 128 * 2 * 2 / 1e9
        ms = a.elapsed_time(b) / 10
        out.append(f"p={p}: {ms * 1000:.0f} us (KV {kv_gb:.2f} GB -> {kv_gb / (ms * 1e-3):.0f} GB/s)")
    print(f"B={B}: " + "; ".join(out), flush=True)
    del dec
