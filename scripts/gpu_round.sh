# GPU round: parity tests, bench, perf probe.
set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
