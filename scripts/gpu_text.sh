#!/bin/bash
# text side of the prefill on the GPU: parity tests, launch list of one call per length, native vs PyTorch glue timing
set -x
timeout 300 python -m pytest tests/test_text_projection.py -m gpu -q -s -p no:cacheprovider > gpurun_out/text_tests.log 2>&1; echo "text pytest rc=$?"
grep -E "^\[|passed|failed|Error|error|bit-equal" gpurun_out/text_tests.log | head -60
cat > /tmp/text_case.py <<'PY'
import sys, json, torch
sys.path.insert(0, "."); sys.path.insert(0, "qwen-megakernel-tts_b200")
import bench
from qwen_megakernel.synthetic import synthetic_tts_weights
w = synthetic_tts_weights(seed=1234, num_layers=1, max_seq_len=64, text_vocab=512, include_talker=False, include_code_predictor=False)
wg = {k: v.cuda() for k, v in w.items() if k.startswith("text_")}
print(json.dumps(bench.time_text_projection(wg, torch.device("cuda", 0), n_tokens=(8, 32, 64, 128, 512))))
PY
timeout 200 python /tmp/text_case.py > gpurun_out/text_time.json 2> gpurun_out/text_time.err; echo "text time rc=$?"; cat gpurun_out/text_time.json; tail -3 gpurun_out/text_time.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:"kt_|qmk_bgemm" -c 60 --csv --log-file gpurun_out/text_launches.csv python /tmp/text_case.py > gpurun_out/text_ncu.log 2>&1; echo "ncu rc=$?"
