import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
from qwen_megakernel.model_tts import CodePredictorKernel
from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to
w = weights_to(synthetic_tts_weights(seed=1234, max_seq_len=2048), "cuda")
cp = CodePredictorKernel(w, device="cuda")
hid = synthetic_inputs(7, 1).float().cuda()[0]
for sample in (False, True):
    for _ in range(5): cp.predict(hid, 100, w["embed_weight"], do_sample=sample, temperature=0.9, top_k=50)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): cp.predict(hid, 100, w["embed_weight"], do_sample=sample, temperature=0.9, top_k=50)
    b.record(); torch.cuda.synchronize()
    print("sampled" if sample else "greedy", round(a.elapsed_time(b) * 10, 1), "us per frame")
codes = [cp.predict(hid, 100, w["embed_weight"], do_sample=True, temperature=0.9, top_k=50).cpu().tolist() for _ in range(3)]
print("codes", codes)
