set -x
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 300 python scripts/perf_probe.py --configs "800,4500;800,4500" --trace > gpurun_out/probe.log 2>&1; echo rc=$?; grep -E "poll_delay|attn  total|per-layer" gpurun_out/probe.log
