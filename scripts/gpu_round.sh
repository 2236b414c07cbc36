set -x
timeout 900 python -m pytest tests/test_gpu_batched.py -m gpu -q -x -s --timeout 600 > gpurun_out/gpu_batched.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/gpu_batched.log
timeout 400 python scripts/batched_probe.py 2>&1 | tail -5
