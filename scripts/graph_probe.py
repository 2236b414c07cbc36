"""Is the batched launch chain bound by the host's launch rate?  Times one batched talker step (a) as issued from the host,
(b) the host's own time to issue it, (c) replayed from a CUDA graph captured from the same C-ABI calls."""
import sys, time, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
from qwen_megakernel.model_tts import BatchedTTSDecoder
from qwen_megakernel.synthetic import synthetic_tts_weights, weights_to

w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=2048), "cuda")
for B in (16, 64):
    dec = BatchedTTSDecoder(w, B, max_seq_len=2048)
    tok = torch.arange(B, dtype=torch.int32, device="cuda") + 5
    for _ in range(5): dec.step(tok)
    torch.cuda.synchronize()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec.reset(); a.record(); t0 = time.perf_counter()
    for _ in range(n): dec.step(tok)
    t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
    print(f"B={B}: stream launches {a.elapsed_time(b) / n * 1000:.0f} us/step on the device, host issue {(t1 - t0) / n * 1e6:.0f} us/step", flush=True)
    # graph
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    dec.reset()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        dec.step(tok)
        torch.cuda.synchronize()
        dec.reset()
        with torch.cuda.graph(g, stream=s):
            toks, hid = dec.step(tok)
    torch.cuda.synchronize()
    ref = BatchedTTSDecoder(w, B, max_seq_len=2048)
    dec.reset(); dec._steps = 0
    ok = True
    for i in range(4):
        g.replay(); torch.cuda.synchronize()
        t_ref, _ = ref.step(tok); torch.cuda.synchronize()
        ok &= bool((t_ref == toks).all())
    print(f"   graph replay == stream launches: {ok}; positions {dec.positions[:3].tolist()}")
    dec.reset()
    for _ in range(5): g.replay()
    torch.cuda.synchronize(); dec.reset()
    a.record()
    for _ in range(n): g.replay()
    b.record(); torch.cuda.synchronize()
    print(f"B={B}: graph replay {a.elapsed_time(b) / n * 1000:.0f} us/step", flush=True)
    del dec, ref, g
