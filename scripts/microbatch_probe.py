"""B = 64 concurrent utterances as one batch of 64 vs two batches of 32 / four of 16 on separate streams (graph replays overlap)."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
from qwen_megakernel.model_tts import BatchedFrameLoop
from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to

w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=2048), "cuda")
total = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for parts in (1, 2, 4):
    B = total // parts
    if B % 16: continue
    loops, streams = [], []
    for i in range(parts):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            loop = BatchedFrameLoop(w, B, max_seq_len=512)
            loop.start(synthetic_inputs(1 + i, 8 * B).cuda().view(8, B, 1024))
            extra = synthetic_inputs(50 + i, B).cuda()
            for _ in range(4): loop.frame(extra)
        loops.append((loop, extra)); streams.append(st)
    torch.cuda.synchronize()
    n = 30
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for st in streams: st.wait_stream(torch.cuda.current_stream())
    for _ in range(n):
        for (loop, extra), st in zip(loops, streams):
            with torch.cuda.stream(st):
                loop.frame(extra)
    for st in streams: torch.cuda.current_stream().wait_stream(st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"{total} streams as {parts} x {B}: {ms:.3f} ms per frame of all streams = {total * 1000 / ms:.0f} codec frames/s", flush=True)
    del loops
