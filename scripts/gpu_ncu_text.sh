#!/bin/bash
# ncu --set full of one text-projection chain (5 kernels, T = 64) after the same program ran clean without ncu
cat > /tmp/text_one.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
from qwen_megakernel.model_tts import TextProjectionKernel
from qwen_megakernel.synthetic import synthetic_tts_weights
w = synthetic_tts_weights(seed=1234, num_layers=1, max_seq_len=64, text_vocab=512, include_talker=False, include_code_predictor=False)
wg = {k: v.cuda() for k, v in w.items() if k.startswith("text_")}
tp = TextProjectionKernel(wg, device="cuda")
ids = torch.randint(0, 512, (64,), device="cuda")
for _ in range(3): out = tp.embed_text_ids(ids)
torch.cuda.synchronize(); print("ok", float(out.float().abs().max()))
PY
timeout 120 python /tmp/text_one.py && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"kt_|qmk_bgemm" -s 10 -c 5 -o gpurun_out/prof_text python /tmp/text_one.py > gpurun_out/ncu_text.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_text.log
