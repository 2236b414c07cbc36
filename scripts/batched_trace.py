#!/usr/bin/env python3
"""Per-phase time of one persistent batched step (CTA 0's barrier stamps): python scripts/batched_trace.py [B]"""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))
os.environ["QMK_BATCHED_TRACE"] = "1"
os.environ["QMK_BATCHED_PERSISTENT"] = "1"
import numpy as np, torch
from qwen_megakernel.model_tts import BatchedTTSDecoder
from qwen_megakernel.synthetic import synthetic_tts_weights, weights_to
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=512), "cuda")
bd = BatchedTTSDecoder(w, B, max_seq_len=512)
tok = torch.full((B,), 2149, dtype=torch.int32, device="cuda")
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    t, _ = bd.step(tok); tok.copy_(t)
buf = (ctypes.c_longlong * 1024)()
n = bd._lib.qmk_batched_trace_read(bd._handle, torch.cuda.current_stream().cuda_stream, buf, 1024)
raw = np.array(buf[:n], dtype=np.float64)
st = raw[:2 * (2 + 28 * 8)].reshape(-1, 2)          # [barrier][enter, leave]
for kind, nm in enumerate(['qkv', 'o', 'gu', 'down']):
    d = raw[600 + kind * 20: 600 + kind * 20 + 16]
    print(nm, 'gemm item stamps rel. to start (A ready, B ready per k-block; 14 = done seen, 15 = epilogue end):', [int(v - d[0]) if v else 0 for v in d])
names = ["qkv gemm", "qkv+attn", "o gemm", "resid+norm", "gu gemm", "swiglu", "down gemm", "resid+norm2"]
L = 28
work = np.zeros(8); wait = np.zeros(8)
for l in range(2, L):
    for k in range(8):
        i = 1 + l * 8 + k
        work[k] += st[i, 0] - st[i - 1, 1]
        wait[k] += st[i, 1] - st[i, 0]
print(f"B={B}: cycles per phase (CTA 0; mean over layers 2..27)   work   barrier-wait")
for k in range(8):
    print(f"  {names[k]:12s} {work[k] / (L - 2):8.0f} {wait[k] / (L - 2):8.0f}")
print(f"  layer total {np.sum(work + wait) / (L - 2):.0f} cycles; step {st[-1, 1] - st[0, 0]:.0f} cycles")
