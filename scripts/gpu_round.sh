# GPU round: parity tests, bench, perf probe.
set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e'], d['roofline']['launch_us'], d['roofline']['frac'], d['cp_frame'])"
