set -x
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['e2e_dropin_loop']['value'], d['roofline']['launch_us'], round(d['roofline']['frac'],3), d['cp_frame']['ms'])"; tail -3 gpurun_out/bench.err
