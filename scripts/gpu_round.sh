# GPU round: parity tests, then the perf probe with the per-phase trace.
set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/gpu_tests.log
timeout 300 python scripts/perf_probe.py --configs "450,2500;300,2500;600,3500;450,4500" --trace > gpurun_out/probe.log 2>&1; echo rc=$?; cat gpurun_out/probe.log | tail -20
