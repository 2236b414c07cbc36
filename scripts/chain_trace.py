"""Timeline of one layer of the batched launch chain (globaltimer stamps of the first / last CTA of every kernel):
python scripts/chain_trace.py [B]"""
import ctypes, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
from qwen_megakernel.model_tts import BatchedTTSDecoder
from qwen_megakernel.synthetic import synthetic_tts_weights, weights_to

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=2048), "cuda")
dec = BatchedTTSDecoder(w, B, max_seq_len=2048)
tok = torch.arange(B, dtype=torch.int32, device="cuda") + 5
for _ in range(20): dec.step(tok)
torch.cuda.synchronize()
lib, st = dec._lib, torch.cuda.current_stream().cuda_stream
lib.qmk_batched_chain_trace(1, st, None, 0)
for _ in range(3): dec.step(tok)
buf = (ctypes.c_ulonglong * 32000)()
n = lib.qmk_batched_chain_trace(0, st, buf, 16000)
rec = sorted((buf[2 * i + 1], buf[2 * i]) for i in range(n) if buf[2 * i + 1])
names = {1: "gemm", 2: "input", 3: "resid_norm", 4: "qkv_attn", 5: "gu_epi", 6: "head", 7: "gu_gemm"}
ev = {0: "entry", 1: "dep ok", 2: "mma done", 3: "exit"}
ev2 = {3: "loads ok", 4: "q ready"}    # what stamp 2 marks in the epilogue kernels
print(f"B={B}: {n} records")
# the middle step, layers 10..11
t_first = rec[0][0]
third = [r for r in rec if r[0] >= rec[len(rec) // 2][0]]
t0 = third[0][0]
for ns, tag in third[:120]:
    print(f"{(ns - t0) / 1000:9.2f} us  {names.get(tag >> 4, tag >> 4):10s} {(ev2.get(tag >> 4, ev[2]) if (tag >> 1) & 7 == 2 else ev[(tag >> 1) & 7]):8s} {'last CTA' if tag & 1 else 'first CTA'}")
