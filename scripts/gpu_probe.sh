set -x
timeout 300 python scripts/perf_probe.py --configs "450,2500" --trace > gpurun_out/probe.log 2>&1; echo rc=$?; cat gpurun_out/probe.log | tail -20
