#!/bin/bash
# round 2: GPU test suite + smoke
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
