#!/usr/bin/env python3
"""GPU probe of the batched multi-stream decode: ms/step and stream-steps/s, eager launches vs one CUDA graph."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))

import torch  # noqa: E402

from qwen_megakernel.model_tts import BatchedTTSDecoder  # noqa: E402
from qwen_megakernel.synthetic import synthetic_tts_weights, weights_to  # noqa: E402

WEIGHT_BYTES = 887_228_928
KV_BYTES_PER_POS = 114_688


def timed(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    torch.cuda.set_device(0)
    S = 512
    w = weights_to(synthetic_tts_weights(max_seq_len=S), "cuda")
    for B in (16, 32, 64):
        bd = BatchedTTSDecoder(w, B, max_seq_len=S)
        tok = torch.full((B,), 2149, dtype=torch.int32, device="cuda")

        def eager():
            t, _ = bd.step(tok)
            tok.copy_(t)

        for _ in range(5):
            eager()
        bd.reset()
        ms_eager = timed(eager, 100)
        bd.reset()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            eager()
        torch.cuda.current_stream().wait_stream(s)
        bd.reset()
        with torch.cuda.graph(g):
            eager()
        bd.reset()
        for _ in range(5):
            g.replay()
        bd.reset()
        ms_graph = timed(g.replay, 200)
        mean_pos = 100
        bytes_step = WEIGHT_BYTES + B * KV_BYTES_PER_POS * (mean_pos + 2)
        print(f"B={B:3d}: eager {ms_eager*1e3:8.1f} us/step ({B/ms_eager*1e3:9.0f} stream-steps/s)   graph {ms_graph*1e3:8.1f} us/step "
              f"({B/ms_graph*1e3:9.0f} stream-steps/s, {bytes_step/ms_graph/1e6:6.0f} GB/s algorithmic)", flush=True)
        del bd, g


if __name__ == "__main__":
    main()
