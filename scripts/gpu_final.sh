#!/bin/bash
# final-tree evidence: text tests, smoke, driver-style bench, then the whole GPU suite with durations
set -x
timeout 300 python -m pytest tests/test_text_projection.py -m gpu -q -s -p no:cacheprovider > gpurun_out/text_tests.log 2>&1; echo "text pytest rc=$?"
grep -E "passed|failed|Error" gpurun_out/text_tests.log | tail -3
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_k20.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_k20.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 4), d["roofline"]["launch_us"])
    print("cp", d["cp_frame"]["ms"], d["cp_frame"]["ms_sampled"], "batched", {k: round(v["ms_per_step"], 3) for k, v in d.get("batched", {}).items()})
    print("prefill", {k: round(v, 3) for k, v in d["prefill"].items() if k != "note"})
    print("text", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["text_projection"].items() if k != "note"})
    print("cpu", d.get("cpu_baseline"), "clocks", d.get("clocks"))
except Exception as e:
    print("bench unreadable", e)
PY
timeout ${SUITE_LIMIT:-720} python -m pytest tests -m gpu -q -p no:cacheprovider --durations=12 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/gpu_tests.log
