# GPU round: parity tests + perf probe with trace.
set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 300 python scripts/perf_probe.py --configs "450,4500" --trace > gpurun_out/probe.log 2>&1; echo rc=$?; tail -20 gpurun_out/probe.log
